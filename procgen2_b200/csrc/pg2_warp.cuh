// Execution context of the level-generation ("reset") kernel: ONE WARP per resetting
// environment. Level generators are long, sequential, data-dependent programs (Kruskal mazes,
// cellular automata, BFS; SURVEY §7 hard part (c)); running them one-thread-per-env would
// serialise 32 different control flows per warp. Instead the 32 lanes of a warp execute the
// same generator for the same environment *redundantly* (identical registers, identical RNG
// position, shared-memory reads are broadcasts) and split only the bulk work — MT19937
// twists, tile-map fills, cellular-automaton passes, copies to HBM — by lane id.
#pragma once
#include "pg2_rng.cuh"
#ifdef PG2_HOSTSIM
#include <stdio.h>
#include <stdlib.h>
#endif

namespace pg2 {

constexpr int RESET_WARPS_PER_CTA = 2;
constexpr int RESET_ARENA_BYTES = 64 * 1024;   // largest per-warp scratch of any game (G::RESET_ARENA is what a game gets)

struct WarpCtx {
    WarpMt rng;
    int lane;
    char* arena;       // per-warp shared-memory scratch
    int arena_off;
    int arena_cap;     // bytes available (G::RESET_ARENA)
    int mode;          // distribution mode (tilemap.h Config of the game): 0 easy, 1 hard, 2 memory / extreme

    template <class T>
    PG2_DEV_NOINLINE T* alloc(int count) {
        int off = (arena_off + 15) & ~15;
        arena_off = off + (int)sizeof(T) * count;
#ifndef PG2_HOSTSIM
        if (arena_off > arena_cap) __trap();   // a game outgrew its G::RESET_ARENA: fail loudly, never corrupt
#else
        if (arena_off > arena_cap) { fprintf(stderr, "reset arena overflow: %d > %d\n", arena_off, arena_cap); abort(); }
        if (getenv("PG2_ARENA_TRACE")) { static int hw = 0; if (arena_off > hw) { hw = arena_off; fprintf(stderr, "arena high water %d\n", hw); } }
#endif
        return (T*)(arena + off);
    }
    template <class T>
    PG2_DEV_NOINLINE void fill(T* p, int count, T v) {
        for (int i = lane; i < count; i += WARP_LANES) p[i] = v;
        __syncwarp();
    }
};

// Ordered, lane-parallel stream compaction: out[0 .. n) = ascending indices i in [0, count) with pred(i). Returns n.
template <class Pred, class T>
PG2_DEV int warp_compact(WarpCtx& w, int count, Pred pred, T* out) {
    int n = 0;
    __syncwarp();
    for (int base = 0; base < count; base += WARP_LANES) {
        int i = base + w.lane;
        bool p = i < count && pred(i);
        uint32_t m = __ballot_sync(0xffffffffu, p);
        if (p) out[n + __popc(m & ((1u << w.lane) - 1u))] = (T)i;
        n += __popc(m);
    }
    __syncwarp();
    return n;
}

}  // namespace pg2
