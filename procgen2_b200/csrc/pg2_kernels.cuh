// Bodies of the three per-step kernels, shared between the __global__ entry points in
// engine.cu and the single-threaded test harness tests/hostsim/ (see pg2_platform.cuh).
#pragma once
#include "pg2_render.cuh"
#include "pg2_state.cuh"
#include "pg2_warp.cuh"

namespace pg2 {

// rng.seed(seed) + process-lifetime defaults (cenv_make)
PG2_DEV_NOINLINE void seed_body(const CommonState& c, int env, uint32_t seed, bool init_persistent) {
    mt_seed(c.mt + (size_t)env * MT_N, seed);
    c.mti[env] = MT_N;
    if (init_persistent) {
        c.cam_x[env] = 0.0f; c.cam_y[env] = 0.0f;   // Renderer::camera_position{0} (renderer.h:18)
        c.sprites_valid[env] = 0; c.ep_steps[env] = 0; c.fault[env] = 0; c.view_valid[env] = 0;
    }
}

// cenv_step for one env (without the render): returns "episode over" (terminated or truncated)
template <class G>
PG2_DEV bool step_body(const typename G::State& s, const CommonState& c, int env, int action, float* reward,
                       uint8_t* terminated, uint8_t* truncated, int max_episode_steps, const StepCtx& ctx) {
    float r = 0.0f;
    bool term = false;
    if (G::LANE_AWARE || ctx.nlanes == 1) {
        term = G::step(s, c, env, action, &r, ctx);
    } else {   // the game's step is single-threaded: the warp's leader runs it, the result is broadcast
        if (ctx.leader()) term = G::step(s, c, env, action, &r, StepCtx{ 0, 1 });
        __syncwarp();
        term = warp_bcast((int)term) != 0;
        r = warp_bcast(r);
    }
    bool trunc = false;
    int ep = c.ep_steps[env] + 1;
    if (max_episode_steps > 0 && ep >= max_episode_steps && !term) trunc = true;
    ctx.sync();
    if (ctx.leader()) {
        c.ep_steps[env] = ep;
        reward[env] = r;
        terminated[env] = term ? 1 : 0;
        truncated[env] = trunc ? 1 : 0;
    }
    return term || trunc;
}

// reset() for one env by one warp; `mt` = 624 words of per-warp scratch, `arena` = G::RESET_ARENA bytes
template <class G>
PG2_DEV void reset_body(const typename G::State& s, const CommonState& c, int env, uint32_t* mt, char* arena, int lane) {
    uint32_t* gmt = c.mt + (size_t)env * MT_N;
    for (int i = lane; i < MT_N; i += WARP_LANES) mt[i] = gmt[i];
    __syncwarp();
    WarpCtx ctx;
    ctx.rng.mt = mt; ctx.rng.idx = c.mti[env]; ctx.rng.lane = lane;
    ctx.lane = lane; ctx.arena = arena; ctx.arena_off = 0; ctx.arena_cap = G::RESET_ARENA;
    ctx.mode = c.mode < 0 ? G::DEFAULT_MODE : c.mode;
    G::regenerate(s, c, env, ctx);
    __syncwarp();
    for (int i = lane; i < MT_N; i += WARP_LANES) gmt[i] = mt[i];
    if (lane == 0) { c.mti[env] = ctx.rng.idx; c.ep_steps[env] = 0; c.view_valid[env] = 0; }   // new level: the cached view is stale
    __syncwarp();
}

// Debug build (-DPG2_PHASE_TIMERS): wall cycles of the render phases of a frame, as seen by thread 0 of the CTA, summed over
// all frames: [0] frames, [1] ticket + frame_begin, [2] build_frame, [3] finalize, [4] rasterise (thread 0's warp).
#ifdef PG2_PHASE_TIMERS
__device__ unsigned long long g_phase[8];
#define PG2_PHASE_MARK(k) do { if (threadIdx.x == 0) { long long now__ = clock64(); atomicAdd(&g_phase[k], (unsigned long long)(now__ - phase_t__)); phase_t__ = now__; } } while (0)
#else
#define PG2_PHASE_MARK(k) do { } while (0)
#endif

template <class G>
using FrameOf = FrameT<G::MAX_POST, G::ROTATES, G::TILE_CLASSES, G::WIN_ROWS, G::BLIT_UNROLL>;

// render_game(true) + RGBA->RGB pack for one env by one CTA (f.tileword filled by frame_init_tiletex before).
// view_cache (G::STATIC_VIEW games, else nullptr): VIEW_CACHE_BYTES per env = the env's base image;
// c.view_valid[env] says whether it is current.
template <class G>
PG2_DEV void render_body(const typename G::State& s, const CommonState& c, int env, FrameOf<G>& f, const TexInfo* __restrict__ tex,
                         const uint32_t* __restrict__ atlas, uint8_t* __restrict__ obs, uint8_t* __restrict__ view_cache = nullptr,
                         bool begin_and_sync = true) {
    uint8_t* cache = (G::STATIC_VIEW && view_cache) ? view_cache + (size_t)env * VIEW_CACHE_BYTES : nullptr;
    if (begin_and_sync) {
        frame_begin(f, cache != nullptr && c.view_valid[env] != 0);
        __syncthreads();
    }
#ifdef PG2_PHASE_TIMERS
    long long phase_t__ = clock64();
#endif
    G::build_frame(s, c, env, f, tex);
    __syncthreads();   // the only barrier between the frame builder's smem writes and their readers
    PG2_PHASE_MARK(2);
    frame_finalize<G>(f);
    PG2_PHASE_MARK(3);
    // a frame that needs the general ordered path everywhere (never observed) is neither cached nor marked
    uint8_t* img = (cache && (f.reuse || !f.wide)) ? cache : nullptr;
    frame_rasterise<G>(f, atlas, obs + (size_t)env * OBS_BYTES, img);
    PG2_PHASE_MARK(4);
    if (img && !f.reuse && threadIdx.x == 0) c.view_valid[env] = 1;   // read by later launches only
    if (f.overflow && threadIdx.x == 0) c.fault[env] |= 8;            // tile window taller than G::WIN_ROWS: the frame is wrong
}

// render_game(false) + the RGB read-out of cenv_render (coinrun.cpp:393-411, 443-470) for ONE env: a frame of any size
// with the window size as camera_size, drawn pixel by pixel along the general ordered path (cold path: the human
// viewer). Every CTA of the launch describes the frame for itself and shades the pixels [first, last) of the
// row-major frame; out = width * height * 3 bytes.
template <class G>
PG2_DEV void render_human_body(const typename G::State& s, const CommonState& c, int env, FrameOf<G>& f, const TexInfo* __restrict__ tex,
                               const uint32_t* __restrict__ atlas, uint8_t* __restrict__ out, int width, int height, int first, int last) {
    frame_begin(f, false);
    __syncthreads();
    if (threadIdx.x == 0) { f.view_w = (float)width; f.view_h = (float)height; f.human = 1; }
    __syncthreads();
    G::build_frame(s, c, env, f, tex);
    __syncthreads();
    frame_finalize<G>(f);
    for (int p = first + (int)threadIdx.x; p < last; p += CTA_THREADS) {
        const uint32_t color = shade_human_pixel<G>(f, atlas, p % width, p / width);
        out[3 * (size_t)p] = (uint8_t)color; out[3 * (size_t)p + 1] = (uint8_t)(color >> 8); out[3 * (size_t)p + 2] = (uint8_t)(color >> 16);
    }
}

}  // namespace pg2
