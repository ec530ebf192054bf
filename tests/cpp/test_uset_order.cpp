// TEST: pg2::USetOrder::order (iteration order of a fresh std::unordered_set<int> computed as <= 9 counting-sort
// regroupings, pg2_roomgen.cuh) == the real libstdc++ container, for every set size the generators can produce
// and random distinct key sequences. Single-lane execution of the same code the warp runs.
#define PG2_HOSTSIM 1
#include <stdio.h>
#include <algorithm>
#include <random>
#include <unordered_set>
#include <vector>
#include "../../procgen2_b200/csrc/pg2_roomgen.cuh"

using namespace pg2;

constexpr int ROOM_CELLS = ROOM_MAX_DIM * ROOM_MAX_DIM;   // 45 x 45: the largest cave (jumper / caveflyer memory mode)

int main() {
    std::mt19937 rng(2024);
    std::vector<char> arena(RESET_ARENA_BYTES);
    static uint32_t mt[MT_N];
    int checks = 0;
    for (int trial = 0; trial < 700; trial++) {
        int n = trial < 120 ? trial : (int)(rng() % (ROOM_CELLS + 1));   // every small size (incl. 0, the rehash edges 13/14, 29/30 ...) + random
        std::vector<int> all(ROOM_CELLS);
        for (int i = 0; i < ROOM_CELLS; i++) all[i] = i;
        std::shuffle(all.begin(), all.end(), rng);
        std::vector<uint16_t> keys(all.begin(), all.begin() + n);
        if (trial % 5 == 0) std::sort(keys.begin(), keys.end());   // BFS-like locally ordered sequences too
        std::unordered_set<int> ref;
        for (int i = 0; i < n; i++) ref.insert(keys[i]);
        std::vector<int> want(ref.begin(), ref.end());

        WarpCtx w;
        w.rng.mt = mt; w.rng.idx = 0; w.rng.lane = 0;
        w.lane = 0; w.arena = arena.data(); w.arena_off = 0; w.arena_cap = RESET_ARENA_BYTES;
        uint16_t* bufa = w.alloc<uint16_t>(ROOM_CELLS + 64);
        uint16_t* bufb = w.alloc<uint16_t>(ROOM_CELLS + 64);
        int* first = w.alloc<int>(ROOM_MAX_BUCKETS);
        int* cnt = w.alloc<int>(ROOM_MAX_BUCKETS);
        const uint16_t* got = USetOrder::order(w, keys.data(), n, bufa, bufb, first, cnt);
        checks++;
        for (int i = 0; i < n; i++)
            if (got[i] != want[i]) { printf("MISMATCH trial %d n %d at %d: %d vs %d\n", trial, n, i, got[i], want[i]); return 1; }
    }
    printf("OK %d checks\n", checks);
    return 0;
}
