// TEST: the bit-row forms of Room_Generator::update (cellular automaton) and ::expand_room (path dilation) in
// pg2_roomgen.cuh against straightforward per-cell restatements of room_generator.cpp:21-36 / 166-207, on random grids.
#define PG2_HOSTSIM 1
#include <stdio.h>
#include <random>
#include <vector>
#include "../../procgen2_b200/csrc/pg2_roomgen.cuh"

using namespace pg2;

int main() {
    std::mt19937 rng(7);
    std::vector<char> arena(RESET_ARENA_BYTES);
    static uint32_t mt[MT_N];
    const int dims[3] = { 40, 20, 45 };   // hard, easy, memory (45: rows that are not a multiple of 4 cells)
    for (int trial = 0; trial < 450; trial++) {
        const int W = dims[trial % 3], H = W;
        WarpCtx w;
        w.rng.mt = mt; w.rng.idx = 0; w.rng.lane = 0;
        w.lane = 0; w.arena = arena.data(); w.arena_off = 0; w.arena_cap = RESET_ARENA_BYTES;
        RoomGen rg;
        rg.init(w, W, H);
        const int density = 20 + trial % 60;
        std::vector<uint8_t> g(W * H);
        for (int i = 0; i < W * H; i++) g[i] = rg.grid[i] = (int)(rng() % 100) < density ? 1 : 0;
        // --- update: wall iff >= 5 walls in the 3x3 neighbourhood (self included, out of bounds = wall)
        std::vector<uint8_t> want(W * H);
        for (int x = 0; x < W; x++)
            for (int y = 0; y < H; y++) {
                int n = 0;
                for (int a = -1; a <= 1; a++)
                    for (int b = -1; b <= 1; b++) {
                        int nx = x + a, ny = y + b;
                        n += (nx < 0 || ny < 0 || nx >= W || ny >= H) ? 1 : g[ny + H * nx];
                    }
                want[y + H * x] = n >= 5;
            }
        rg.update(w);
        for (int i = 0; i < W * H; i++)
            if (rg.grid[i] != want[i]) { printf("MISMATCH update trial %d cell %d\n", trial, i); return 1; }
        // --- expand: `rounds` dilations of a seed set over space cells (8-neighbourhood, members on walls do not spread)
        g.assign(rg.grid, rg.grid + W * H);
        std::vector<uint16_t> seeds;
        for (int k = 0; k < 1 + (int)(rng() % 40); k++) seeds.push_back((uint16_t)(rng() % (W * H)));
        std::vector<uint8_t> mem(W * H, 0), nxt(W * H);
        for (uint16_t c : seeds) mem[c] = 1;
        const int rounds = 1 + trial % 5;
        for (int r = 0; r < rounds; r++) {
            for (int i = 0; i < W * H; i++) {
                int v = mem[i];
                if (!v && g[i] == 0) {
                    int x = i / H, y = i % H;
                    for (int a = -1; a <= 1 && !v; a++)
                        for (int b = -1; b <= 1 && !v; b++) {
                            int nx = x + a, ny = y + b;
                            if ((a || b) && nx >= 0 && ny >= 0 && nx < W && ny < H && mem[ny + H * nx] && g[ny + H * nx] == 0) v = 1;
                        }
                }
                nxt[i] = (uint8_t)v;
            }
            mem = nxt;
        }
        for (size_t k = 0; k < seeds.size(); k++) rg.path[k] = seeds[k];
        rg.expand(w, rg.path, (int)seeds.size(), rounds, rg.mark);
        for (int i = 0; i < W * H; i++)
            if (rg.mark[i] != mem[i]) { printf("MISMATCH expand trial %d cell %d\n", trial, i); return 1; }
    }
    printf("OK\n");
    return 0;
}
