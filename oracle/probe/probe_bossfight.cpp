/* TEST INFRASTRUCTURE — see probe_common.h */
#include "common_systems.h"
#include "probe/probe_common.h"
#include <memory>
extern std::shared_ptr<System_Agent> agent;
extern std::shared_ptr<System_Mob_AI> mob_ai;
extern "C" {
/* System_Mob_AI::Config::mode (common_systems.h:61-65; 0 easy, 1 hard); the system exists after cenv_make, so call it then */
void pg2o_set_mode(int mode) { mob_ai->config.mode = (Distribution_Mode)mode; }
void pg2o_tile_dims(int* wh) { wh[0] = 0; wh[1] = 0; }
void pg2o_tiles(int32_t*) {}
int pg2o_floats(float* out, int cap) {
    int n = 0;
    for (auto const& e : agent->entities) {
        auto& t = c.get_component<Component_Transform>(e);
        out[n++] = t.position.x; out[n++] = t.position.y;
    }
    return n;
}
}
