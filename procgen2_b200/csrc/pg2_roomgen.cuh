// Cave generator shared by caveflyer and jumper — device restatement of Room_Generator
// (games/caveflyer/room_generator.cpp == games/jumper/room_generator.cpp, byte-identical):
//   update()          :21-36   cellular automaton, Moore neighbourhood incl. self, out of bounds = wall
//   build_room()      :38-78   BFS flood fill into an unordered_set (start cell only enters when re-discovered, Q19)
//   find_best_room()  :143-164 first strictly largest room; its unordered_set ITERATION ORDER is observable,
//                              because the callers turn it into the free_cells vector (caveflyer/tilemap.cpp:158-161,
//                              jumper/tilemap.cpp:150-153) — emulated with pg2::USet
//   find_path()       :80-141  BFS with parent links, `covered` omits the source (Q19)
//   expand_room()     :166-207 4 rounds of 8-neighbour dilation restricted to space cells (order-insensitive)
// Bulk passes (automaton, dilation) are spread over the 32 lanes of the warp; the order-sensitive
// searches run on lane 0 and publish their result through shared memory.
#pragma once
#include "pg2_uset.cuh"
#include "pg2_warp.cuh"

namespace pg2 {

constexpr int ROOM_DIM = 40, ROOM_CELLS = ROOM_DIM * ROOM_DIM;
using RoomSet = USet<ROOM_CELLS, 2400>;

// Atomically claim a byte flag (0 -> 1); true for the one thread that claimed it.
PG2_DEV bool claim_cell(uint8_t* p) {
#ifdef PG2_HOSTSIM
    if (*p) return false;
    *p = 1;
    return true;
#else
    uint32_t* word = (uint32_t*)((size_t)p & ~(size_t)3);
    uint32_t bit = 1u << (8u * (uint32_t)((size_t)p & 3));
    return (atomicOr(word, bit) & (0xffu << (8u * (uint32_t)((size_t)p & 3)))) == 0u;
#endif
}

struct RoomGen {
    int W, H;
    uint8_t* grid;        // [y + H * x]: 1 wall, 0 space
    uint8_t* tmp;         // double buffer / membership scratch
    uint8_t* mark;        // visited / in-room / covered
    uint16_t* queue;      // BFS queue, `expanded` of find_path
    uint16_t* parents;
    uint16_t* order;      // best_room in unordered_set iteration order
    uint16_t* path;       // goal_path (src .. dst)
    int* res;             // lane 0 -> all lanes: [0] best_room size, [1] path length
    RoomSet* set;

    PG2_DEV_NOINLINE void init(WarpCtx& w, int width, int height) {
        W = width; H = height;
        grid = w.alloc<uint8_t>(ROOM_CELLS);
        tmp = w.alloc<uint8_t>(ROOM_CELLS);
        mark = w.alloc<uint8_t>(ROOM_CELLS);
        queue = w.alloc<uint16_t>(ROOM_CELLS + 64);
        parents = w.alloc<uint16_t>(ROOM_CELLS + 64);
        order = w.alloc<uint16_t>(ROOM_CELLS);
        path = w.alloc<uint16_t>(ROOM_CELLS + 64);
        res = w.alloc<int>(4);
        set = w.alloc<RoomSet>(1);
    }

    PG2_DEV int get(int x, int y) const { return (x < 0 || y < 0 || x >= W || y >= H) ? 1 : grid[y + H * x]; }

    // Room_Generator::update
    PG2_DEV_NOINLINE void update(WarpCtx& w) {
        __syncwarp();
        for (int i = w.lane; i < W * H; i += WARP_LANES) {
            int x = i / H, y = i % H, n = 0;
            for (int a = -1; a <= 1; a++)
                for (int b = -1; b <= 1; b++) n += get(x + a, y + b) == 1;
            tmp[i] = n >= 5 ? 1 : 0;
        }
        __syncwarp();
        for (int i = w.lane; i < W * H; i += WARP_LANES) grid[i] = tmp[i];
        __syncwarp();
    }

    // Room_Generator::find_best_room -> order[0 .. n) = iteration order of best_room; returns n
    PG2_DEV_NOINLINE int find_best_room(WarpCtx& w) {
        // pass 1 (lane-parallel, order-insensitive): size of every room in scan order. A one-cell room yields an
        // EMPTY set (its start cell is never re-discovered); the first room always replaces the initial
        // best_room_size of -1.
        const int cells = W * H;
        __syncwarp();
        for (int i = w.lane; i < cells; i += WARP_LANES) mark[i] = 0;
        __syncwarp();
        int best_start = -1, best_size = -1;
        for (int i = 0; i < cells; i++) {
            if (grid[i] != 0 || mark[i]) continue;          // uniform: every lane reads the same cell
            __syncwarp();
            if (w.lane == 0) { mark[i] = 1; queue[0] = (uint16_t)i; res[2] = 1; }
            __syncwarp();
            int head = 0, tail = 1;
            while (head < tail) {                            // one BFS level chunk per iteration
                for (int q = head + w.lane; q < tail; q += WARP_LANES) {
                    int cur = queue[q], x = cur / H, y = cur % H;
                    const int nx[4] = { x - 1, x, x, x + 1 }, ny[4] = { y, y - 1, y + 1, y };
                    for (int k = 0; k < 4; k++) {
                        if (nx[k] < 0 || ny[k] < 0 || nx[k] >= W || ny[k] >= H) continue;
                        int nxt = ny[k] + H * nx[k];
                        if (grid[nxt] == 0 && claim_cell(&mark[nxt])) queue[atomicAdd(&res[2], 1)] = (uint16_t)nxt;
                    }
                }
                __syncwarp();
                head = tail; tail = res[2];
                __syncwarp();
            }
            int size = tail >= 2 ? tail : 0;
            if (size > best_size) { best_size = size; best_start = i; }
        }
        __syncwarp();
        for (int i = w.lane; i < cells; i += WARP_LANES) mark[i] = 0;
        __syncwarp();
        if (w.lane == 0) {
            // pass 2 (ordered): the winning room again, exactly as build_room inserts it
            int n = 0;
            if (best_size > 0) {
                set->init(1);
                int head = 0, tail = 0;
                queue[tail++] = (uint16_t)best_start;
                while (head < tail) {
                    int cur = queue[head++], x = cur / H, y = cur % H;
                    const int nx[4] = { x - 1, x, x, x + 1 }, ny[4] = { y, y - 1, y + 1, y };
                    for (int k = 0; k < 4; k++) {
                        if (nx[k] < 0 || ny[k] < 0 || nx[k] >= W || ny[k] >= H) continue;
                        int nxt = ny[k] + H * nx[k];
                        if (!mark[nxt] && grid[nxt] == 0) { mark[nxt] = 1; queue[tail++] = (uint16_t)nxt; set->insert_new(nxt); }
                    }
                }
                n = set->order(order);
            }
            res[0] = n;
        }
        __syncwarp();
        return res[0];
    }

    // Room_Generator::find_path -> path[0 .. len) from src to dst; returns len (0: none)
    PG2_DEV_NOINLINE int find_path(WarpCtx& w, int src, int dst) {
        __syncwarp();
        for (int i = w.lane; i < W * H; i += WARP_LANES) mark[i] = 0;   // `covered` (does not contain src)
        __syncwarp();
        if (w.lane == 0) {
            int len = 0;
            if (grid[src] == 0) {
                int count = 0, search = 0;
                queue[count] = (uint16_t)src; parents[count] = 0xffff; count++;
                while (search < count) {
                    int cur = queue[search];
                    if (cur == dst) break;
                    int x = cur / H, y = cur % H;
                    const int nx[4] = { x - 1, x, x, x + 1 }, ny[4] = { y, y - 1, y + 1, y };
                    for (int k = 0; k < 4; k++) {
                        if (nx[k] < 0 || ny[k] < 0 || nx[k] >= W || ny[k] >= H) continue;
                        int nxt = ny[k] + H * nx[k];
                        if (!mark[nxt] && grid[nxt] == 0) {
                            queue[count] = (uint16_t)nxt; parents[count] = (uint16_t)search; count++;
                            mark[nxt] = 1;
                        }
                    }
                    search++;
                }
                if (search < count && queue[search] == dst) {
                    int k = search;
                    while (k != 0xffff) { len++; k = parents[k]; }
                    k = search;
                    for (int j = len - 1; j >= 0; j--) { path[j] = queue[k]; k = parents[k]; }
                }
            }
            res[1] = len;
        }
        __syncwarp();
        return res[1];
    }

    // wide_path = goal_path dilated `rounds` times (Room_Generator::expand_room); member[] is the result
    PG2_DEV_NOINLINE void expand(WarpCtx& w, const uint16_t* cells_in, int n, int rounds, uint8_t* member) {
        __syncwarp();
        for (int i = w.lane; i < W * H; i += WARP_LANES) member[i] = 0;
        __syncwarp();
        for (int i = w.lane; i < n; i += WARP_LANES) member[cells_in[i]] = 1;
        __syncwarp();
        for (int r = 0; r < rounds; r++) {
            for (int i = w.lane; i < W * H; i += WARP_LANES) {
                int v = member[i];
                if (!v && grid[i] == 0) {
                    int x = i / H, y = i % H;
                    for (int a = -1; a <= 1 && !v; a++)
                        for (int b = -1; b <= 1; b++) {
                            int nx = x + a, ny = y + b;
                            if ((a || b) && nx >= 0 && ny >= 0 && nx < W && ny < H) {
                                int j = ny + H * nx;
                                if (member[j] && grid[j] == 0) { v = 1; break; }
                            }
                        }
                }
                tmp[i] = (uint8_t)v;
            }
            __syncwarp();
            for (int i = w.lane; i < W * H; i += WARP_LANES) member[i] = tmp[i];
            __syncwarp();
        }
    }
};

}  // namespace pg2
