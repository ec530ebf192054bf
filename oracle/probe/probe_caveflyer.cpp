/* TEST INFRASTRUCTURE — see probe_common.h */
#include "tilemap.h"
#include "common_systems.h"
#include "probe/probe_common.h"
#include <memory>
extern std::shared_ptr<System_Tilemap> tilemap;
extern std::shared_ptr<System_Agent> agent;
extern System_Tilemap::Config tilemap_config;   /* games/caveflyer/caveflyer.cpp: the generator's compile-time Config, a global */
extern "C" {
/* A probe that WRITES: selects a distribution mode the reference only offers at compile time (tilemap.h Config::mode) —
 * call before cenv_make. */
void pg2o_set_mode(int mode) { tilemap_config.mode = (Distribution_Mode)mode; }
void pg2o_tile_dims(int* wh) { wh[0] = tilemap->get_width(); wh[1] = tilemap->get_height(); }
/* out[x * h + y] = tile id at map position (x, y) — the reference's own column-major order */
void pg2o_tiles(int32_t* out) {
    int w = tilemap->get_width(), h = tilemap->get_height();
    for (int x = 0; x < w; x++) for (int y = 0; y < h; y++) out[x * h + y] = (int32_t)tilemap->get(x, y);
}
/* agent transform (+ dynamics when present), then camera */
int pg2o_floats(float* out, int cap) {
    int n = 0;
    for (auto const& e : agent->entities) {
        auto& t = c.get_component<Component_Transform>(e);
        out[n++] = t.position.x; out[n++] = t.position.y;
        if (c.entity_manager.get_signature(e)[c.get_component_type<Component_Dynamics>()]) {
            auto& d = c.get_component<Component_Dynamics>(e);
            out[n++] = d.velocity.x; out[n++] = d.velocity.y;
        }
    }
    out[n++] = gr.camera_position.x; out[n++] = gr.camera_position.y;
    return n;
}
}
