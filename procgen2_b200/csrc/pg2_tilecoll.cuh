// System_Tilemap::get_collision (games/coinrun/tilemap.cpp:323-396; the other games' copies,
// e.g. games/caveflyer/tilemap.cpp:305-366, are this function without the `down_only` branch):
// two passes over the tiles under the rectangle — resolve in y where the overlap is wider than
// tall, then in x otherwise — with the rectangle mutated in place between tiles.
#pragma once
#include "pg2_common.cuh"

namespace pg2 {

enum CollisionType { COLL_NONE = 0, COLL_FULL = 1, COLL_DOWN_ONLY = 2 };

struct CollisionResult { float x, y; bool collided; };

// TileAt(x, y) -> tile id with y already in render space (the callee flips: get(x, H-1-y)).
// `uni` != nullptr: the call is made by ALL lanes of the warp that owns the environment with identical arguments
// (the agent of a lane-aware game); the window's tiles are then fetched one per lane and OR-reduced.
template <class TileAt, class TypeOf>
PG2_DEV CollisionResult tile_collision(Rect rectangle, TileAt tile_at, TypeOf type_of, bool fallthrough = false, float step_y = 0.0f,
                                       const StepCtx* uni = nullptr) {
    bool collided = false;
    const int lower_x = f2i(floorf(rectangle.x));
    const int lower_y = f2i(floorf(rectangle.y));
    const int upper_x = f2i(ceilf(__fadd_rn(rectangle.x, rectangle.w)));
    const int upper_y = f2i(ceilf(__fadd_rn(rectangle.y, rectangle.h)));
    const float center_x = __fadd_rn(rectangle.x, __fmul_rn(rectangle.w, 0.5f));
    const float center_y = __fadd_rn(rectangle.y, __fmul_rn(rectangle.h, 0.5f));
    Rect tile; tile.w = 1.0f; tile.h = 1.0f;
    const int nxw = upper_x - lower_x + 1, nyw = upper_y - lower_y + 1;

    if (nxw <= 4 && nyw <= 4) {
        // Rectangles are at most one tile wide / tall, so the window is at most 3 x 3 (4 x 4 handled): fetch the
        // collision type of every window tile ONCE (independent loads), 2 bits per cell in y-major / x-minor order,
        // then let both passes walk only the non-empty cells in exactly the reference's iteration order.
        uint32_t mask = 0u;
        if (uni != nullptr && uni->nlanes == 32) {
            const int l = uni->lane, ix = l & 3, iy = (l >> 2) & 3;
            if (l < 16 && ix < nxw && iy < nyw) mask = (uint32_t)type_of(tile_at(lower_x + ix, lower_y + iy)) << (2 * l);
            mask = warp_or(mask);
        } else {
            for (int iy = 0; iy < nyw; iy++)
                for (int ix = 0; ix < nxw; ix++)
                    mask |= (uint32_t)type_of(tile_at(lower_x + ix, lower_y + iy)) << (2 * (iy * 4 + ix));
        }
        if (mask == 0u) { CollisionResult r0; r0.x = rectangle.x; r0.y = rectangle.y; r0.collided = false; return r0; }
        for (uint32_t m = mask; m;) {
            const int cell = (__ffs(m) - 1) >> 1;
            const int type = (int)((m >> (2 * cell)) & 3u);
            m &= ~(3u << (2 * cell));
            tile.x = (float)(lower_x + (cell & 3)); tile.y = (float)(lower_y + (cell >> 2));
            Rect col = get_collision_overlap(rectangle, tile);
            if (col.w != 0.0f || col.h != 0.0f) {
                float ccy = __fadd_rn(col.y, __fmul_rn(col.h, 0.5f));
                if (col.w > col.h) {
                    if (type == COLL_DOWN_ONLY) {
                        bool inside = __fsub_rn(__fadd_rn(rectangle.y, rectangle.h), step_y) > tile.y;
                        if (step_y > 0.01f && !fallthrough && !inside) {
                            rectangle.y = ccy > center_y ? __fsub_rn(tile.y, rectangle.h) : __fadd_rn(tile.y, tile.h);
                            collided = true;
                        }
                    } else {
                        rectangle.y = ccy > center_y ? __fsub_rn(tile.y, rectangle.h) : __fadd_rn(tile.y, tile.h);
                        collided = true;
                    }
                }
            }
        }
        for (uint32_t m = mask; m;) {
            const int cell = (__ffs(m) - 1) >> 1;
            const int type = (int)((m >> (2 * cell)) & 3u);
            m &= ~(3u << (2 * cell));
            tile.x = (float)(lower_x + (cell & 3)); tile.y = (float)(lower_y + (cell >> 2));
            Rect col = get_collision_overlap(rectangle, tile);
            if (col.w != 0.0f || col.h != 0.0f) {
                float ccx = __fadd_rn(col.x, __fmul_rn(col.w, 0.5f));
                if (col.w <= col.h && type != COLL_DOWN_ONLY) {
                    rectangle.x = ccx > center_x ? __fsub_rn(tile.x, rectangle.w) : __fadd_rn(tile.x, tile.w);
                    collided = true;
                }
            }
        }
        CollisionResult rr; rr.x = rectangle.x; rr.y = rectangle.y; rr.collided = collided;
        return rr;
    }


    for (int y = lower_y; y <= upper_y; y++)
        for (int x = lower_x; x <= upper_x; x++) {
            int type = type_of(tile_at(x, y));
            if (type == COLL_NONE) continue;
            tile.x = (float)x; tile.y = (float)y;
            Rect col = get_collision_overlap(rectangle, tile);
            if (col.w != 0.0f || col.h != 0.0f) {
                float ccy = __fadd_rn(col.y, __fmul_rn(col.h, 0.5f));
                if (col.w > col.h) {
                    if (type == COLL_DOWN_ONLY) {
                        bool inside = __fsub_rn(__fadd_rn(rectangle.y, rectangle.h), step_y) > tile.y;
                        if (step_y > 0.01f && !fallthrough && !inside) {
                            rectangle.y = ccy > center_y ? __fsub_rn(tile.y, rectangle.h) : __fadd_rn(tile.y, tile.h);
                            collided = true;
                        }
                    } else {
                        rectangle.y = ccy > center_y ? __fsub_rn(tile.y, rectangle.h) : __fadd_rn(tile.y, tile.h);
                        collided = true;
                    }
                }
            }
        }
    for (int y = lower_y; y <= upper_y; y++)
        for (int x = lower_x; x <= upper_x; x++) {
            int type = type_of(tile_at(x, y));
            if (type == COLL_NONE) continue;
            tile.x = (float)x; tile.y = (float)y;
            Rect col = get_collision_overlap(rectangle, tile);
            if (col.w != 0.0f || col.h != 0.0f) {
                float ccx = __fadd_rn(col.x, __fmul_rn(col.w, 0.5f));
                if (col.w <= col.h) {
                    if (type != COLL_DOWN_ONLY) {
                        rectangle.x = ccx > center_x ? __fsub_rn(tile.x, rectangle.w) : __fadd_rn(tile.x, tile.w);
                        collided = true;
                    }
                }
            }
        }
    CollisionResult r; r.x = rectangle.x; r.y = rectangle.y; r.collided = collided;
    return r;
}

}  // namespace pg2
