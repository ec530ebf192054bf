// Observation rasteriser: one CTA renders one environment's 64x64x3 frame.
//
// Replaces render_game() (games/coinrun/coinrun.cpp:443-470 and its six siblings) together with
// the SDL3 software blits it issues (SDL_RenderTextureRotated, renderer.cpp:78/97) and the
// RGBA->RGB pack loop (coinrun.cpp:377-388). Draw order is the reference's painter's order:
//   clear(0,0,0) -> "pre" blits (background) -> tile layer (y-major, x-minor; tilemap.cpp:303-320)
//   -> "post" blits (particles, sprites, agent, HUD) in submission order.
// Instead of executing blits one after another over a framebuffer, every output pixel gathers
// the layers that cover it, in that order, and blends them in registers; the finished frame is
// staged in shared memory and leaves the SM as ONE 12 288-byte TMA bulk store
// (cp.async.bulk.global.shared::cta), i.e. fully coalesced 128-bit+ writes.
//
// The tile layer exploits that render_texture() is separable: a tile's destination columns only
// depend on its x index and its rows only on its y index, so a frame needs <= 32 column and <= 32
// row descriptors instead of up to 27x27 blit records.
#pragma once
#include "pg2_common.cuh"

namespace pg2 {

constexpr int MAX_WIN = 32;        // tile window extent per axis (maze: 27)
constexpr int MAX_PRE = 2;
constexpr int MAX_POST = 192;      // visible post blits (coinrun: 10 particles per visible mob)
#ifndef PG2_RENDER_THREADS
#define PG2_RENDER_THREADS 128
#endif
constexpr int RENDER_THREADS = PG2_RENDER_THREADS;
constexpr uint8_t NO_TILE = 0xff;

// std::sort permutation table. System_Sprite_Render::update sorts (z, entity) pairs
// by z with std::sort (common_systems.cpp:36-38); every sprite of a game has the same z, so the
// comparator is always false and the resulting permutation depends on n only. It is computed on
// the host with the real std::sort (sort_perm.h) and uploaded: sorted[k] = input[perm[n][k]].
constexpr int SORT_MAXN = 128;
#ifdef PG2_HOSTSIM
static const uint8_t* g_sort_perm = nullptr;
#else
__device__ const uint8_t* g_sort_perm;
#endif
PG2_DEV int sort_perm(int n, int k) { return (n <= 16 || n > SORT_MAXN) ? k : g_sort_perm[n * SORT_MAXN + k]; }

struct BlitRot { double s, c; };   // sin/cos of the blit angle (deterministic, see sincos_deg)

// Per-pixel form of a blit (what the frame keeps): coverage test = two unsigned compares, sampling = one
// multiply-add + shift per axis (a horizontal flip is folded into hx / incx, which wrap modulo 2^32 to the
// exact non-negative value), texel address = base + sy * tex_w + sx.
struct alignas(16) FastBlit {
    int16_t x0, y0; uint16_t w, h;              // integer destination rect (SDL truncates the float rect)
    uint32_t hx, incx;                          // sx = (hx + i * incx) >> 16
    uint32_t hy, incy, base;                    // base = tex_offset + s0y * tex_w + s0x
    uint16_t tex_w; uint8_t flags, alpha_mod;   // flags: 1 blend, 2 rotated, 4 invisible
};

struct TileTex { uint32_t offset; uint16_t w; uint8_t blend, cls; };   // tile textures: ids < MAX_TILE_TEX
constexpr int MAX_TILE_TEX = 32;

// Frame description of ONE environment, in shared memory. MAXP = capacity of the post-blit list (per game),
// ROT = whether the game ever rotates a blit (bossfight, caveflyer, jumper HUD).
template <int MAXP, bool ROT>
struct FrameT {
    static constexpr int MAX_POST = MAXP, WORDS = (MAXP + 31) / 32, NROT = ROT ? MAXP : 1;
    static constexpr bool ROTATES = ROT;
    Blit pre[MAX_PRE];
    FastBlit fpre[MAX_PRE];
    FastBlit fpost[MAXP];
    BlitRot post_rot[NROT];
    int npre, npost;
    // tile layer: window origin (tile coordinates, y in render space), extents, descriptors per
    // texture shape class (textures of one class share width and height)
    int tx0, ty0, ncol, nrow, nclass;
    Axis col[2][MAX_WIN];
    Axis row[2][MAX_WIN];
    uint8_t tile_tex[MAX_WIN * MAX_WIN];        // tile texture id per window cell or NO_TILE
    uint8_t col_lo[OBS_W], col_hi[OBS_W];       // per screen column: range of tile columns covering it
    uint8_t row_lo[OBS_H], row_hi[OBS_H];       // (lo > hi: none)
    // for every screen column / row and class: source texel index under the (at most two) covering tiles
    int16_t col_sx[2][2][OBS_W];                // [class][candidate][X] -> source x, -1: not covered
    int16_t row_sy[2][2][OBS_H];
    int32_t pre_sx[MAX_PRE][OBS_W];             // pre blits (backgrounds), resolved per column / row:
    int32_t pre_row[MAX_PRE][OBS_H];            //   texel index = pre_row[k][Y] + pre_sx[k][X], -1: not covered
    int wcount[RENDER_THREADS / 32];            // emit_post_blits: visible blits per warp
    int next_block;                             // frame_rasterise: dynamic hand-out of the 128 pixel blocks to warps
    int wide;                                   // some column / row is covered by more than two tiles (never observed)
    // post-blit binning: 8x4-pixel blocks (one warp's pixels) x MAXP blits
    uint32_t bin[128][WORDS];
    uint8_t bin_any[128];
    TileTex tiletex[MAX_TILE_TEX];              // per CTA (filled once): atlas offset / stride / blend / class
    alignas(16) uint8_t rgb[OBS_BYTES];         // staged output frame
};

// Deterministic sin/cos in degrees, mirrored operation by operation from oracle/raster.c
// (pg2o_sincos_deg): IEEE double add/mul only, fixed order, no FMA.
PG2_DEV_NOINLINE void sincos_deg(double deg, double* s, double* c) {
    double r = fmod(deg, 360.0);
    if (r < 0.0) r = __dadd_rn(r, 360.0);
    int q = (int)__ddiv_rn(__dadd_rn(r, 45.0), 90.0);
    double t = __dsub_rn(r, __dmul_rn((double)q, 90.0));
    double x = __dmul_rn(t, 0.017453292519943295);
    double x2 = __dmul_rn(x, x);
    double ps = -1.0 / 355687428096000.0;
    ps = __dadd_rn(__dmul_rn(ps, x2), 1.0 / 1307674368000.0);
    ps = __dsub_rn(__dmul_rn(ps, x2), 1.0 / 6227020800.0);
    ps = __dadd_rn(__dmul_rn(ps, x2), 1.0 / 39916800.0);
    ps = __dsub_rn(__dmul_rn(ps, x2), 1.0 / 362880.0);
    ps = __dadd_rn(__dmul_rn(ps, x2), 1.0 / 5040.0);
    ps = __dsub_rn(__dmul_rn(ps, x2), 1.0 / 120.0);
    ps = __dsub_rn(__dmul_rn(ps, x2), 1.0 / 6.0);
    double s0 = __dadd_rn(x, __dmul_rn(x, __dmul_rn(x2, ps)));
    double pc = 1.0 / 20922789888000.0;
    pc = __dsub_rn(__dmul_rn(pc, x2), 1.0 / 87178291200.0);
    pc = __dadd_rn(__dmul_rn(pc, x2), 1.0 / 479001600.0);
    pc = __dsub_rn(__dmul_rn(pc, x2), 1.0 / 3628800.0);
    pc = __dadd_rn(__dmul_rn(pc, x2), 1.0 / 40320.0);
    pc = __dsub_rn(__dmul_rn(pc, x2), 1.0 / 720.0);
    pc = __dadd_rn(__dmul_rn(pc, x2), 1.0 / 24.0);
    pc = __dsub_rn(__dmul_rn(pc, x2), 0.5);
    double c0 = __dadd_rn(1.0, __dmul_rn(x2, pc));
    switch (q & 3) {
    case 0: *s = s0;  *c = c0;  break;
    case 1: *s = c0;  *c = -s0; break;
    case 2: *s = -s0; *c = -c0; break;
    default: *s = -c0; *c = s0; break;
    }
}

// ---- blit construction helpers (called by the per-game frame builders) --------------------

// Work item k of a frame builder runs on the first lane of warp k (any thread when simulated).
PG2_DEV bool is_role(int k) { return (int)threadIdx.x == (k * 32) % (int)blockDim.x; }

struct Camera { float x, y, scale; };   // gr.camera_position / gr.camera_scale; camera_size is 64x64

// Renderer::render_texture (renderer.cpp:5-82)
PG2_DEV Blit make_blit(const TexInfo* tex, int tex_id, float px, float py, const Camera& cam,
                                          float scale, float alpha = 1.0f, bool flip_h = false) {
    TexInfo t = tex[tex_id];
    Blit b;
    b.ax = make_axis(px, cam.x, cam.scale, 64.0f, t.w, scale, flip_h, false);
    b.ay = make_axis(py, cam.y, cam.scale, 64.0f, t.h, scale, false, true);
    if (!b.ay.visible) b.ax.visible = 0;
    b.tex_offset = t.offset; b.tex_w = t.w; b.blend = (uint8_t)t.blend;
    // `SDL_SetTextureAlphaMod(tex, 255 * alpha)`: float -> Uint8 truncation (renderer.cpp:57)
    b.alpha_mod = (alpha != 1.0f) ? (uint8_t)f2i(__fmul_rn(255.0f, alpha)) : (uint8_t)255;
    b.flip_h = flip_h; b.rotated = 0;
    return b;
}

// Renderer::render_texture_rotated (renderer.cpp:84-101): whole texture, no culling / cropping,
// angle = rotation * 180.0f / M_PI evaluated in double.
PG2_DEV Blit make_blit_rotated(const TexInfo* tex, int tex_id, float px, float py, const Camera& cam,
                                                  float rotation, float scale, float alpha, BlitRot* rot) {
    TexInfo t = tex[tex_id];
    Blit b;
    float dx = __fadd_rn(__fmul_rn(__fsub_rn(px, cam.x), cam.scale), 32.0f);
    float dy = __fadd_rn(__fmul_rn(__fsub_rn(py, cam.y), cam.scale), 32.0f);
    float dw = __fmul_rn(__fmul_rn((float)t.w, scale), cam.scale);
    float dh = __fmul_rn(__fmul_rn((float)t.h, scale), cam.scale);
    b.ax = make_axis_direct(dx, dw, t.w);
    b.ay = make_axis_direct(dy, dh, t.h);
    if (!b.ay.visible) b.ax.visible = 0;
    b.tex_offset = t.offset; b.tex_w = t.w; b.blend = (uint8_t)t.blend;
    b.alpha_mod = (alpha != 1.0f) ? (uint8_t)f2i(__fmul_rn(255.0f, alpha)) : (uint8_t)255;
    b.flip_h = 0;
    double angle = __ddiv_rn((double)__fmul_rn(rotation, 180.0f), 3.14159265358979323846);
    b.rotated = (angle != 0.0) ? 1 : 0;
    rot->s = 0.0; rot->c = 1.0;
    if (b.rotated) sincos_deg(angle, &rot->s, &rot->c);
    return b;
}

// Blit with an explicit float destination rect and angle in degrees (jumper HUD, jumper.cpp:487-508)
PG2_DEV Blit make_blit_rect(const TexInfo* tex, int tex_id, float dx, float dy, float dw, float dh,
                                               double angle_deg, BlitRot* rot) {
    TexInfo t = tex[tex_id];
    Blit b;
    b.ax = make_axis_direct(dx, dw, t.w);
    b.ay = make_axis_direct(dy, dh, t.h);
    if (!b.ay.visible) b.ax.visible = 0;
    b.tex_offset = t.offset; b.tex_w = t.w; b.blend = (uint8_t)t.blend;
    b.alpha_mod = 255; b.flip_h = 0;
    b.rotated = (angle_deg != 0.0) ? 1 : 0;
    rot->s = 0.0; rot->c = 1.0;
    if (b.rotated) sincos_deg(angle_deg, &rot->s, &rot->c);
    return b;
}

// Wait until every bulk store issued by this thread has finished READING shared memory (see frame_store).
PG2_DEV void frame_store_wait() {
#ifndef PG2_HOSTSIM
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
}

PG2_DEV FastBlit make_fast(const Blit& b) {
    FastBlit fb;
    bool off = b.ax.d0 < -32768 || b.ax.d0 > 32767 || b.ay.d0 < -32768 || b.ay.d0 > 32767 || b.ax.dlen > 65535 || b.ay.dlen > 65535;
    fb.x0 = (int16_t)b.ax.d0; fb.y0 = (int16_t)b.ay.d0;
    fb.w = (uint16_t)max(0, b.ax.dlen); fb.h = (uint16_t)max(0, b.ay.dlen);
    fb.incx = b.flip_h ? 0u - b.ax.inc : b.ax.inc;
    fb.hx = b.flip_h ? b.ax.inc / 2u + (uint32_t)(b.ax.dlen - 1) * b.ax.inc : b.ax.inc / 2u;
    fb.hy = b.ay.inc / 2u; fb.incy = b.ay.inc;
    fb.base = b.tex_offset + (uint32_t)b.ay.s0 * b.tex_w + (uint32_t)b.ax.s0;
    fb.tex_w = b.tex_w;
    fb.alpha_mod = b.alpha_mod;
    // a destination rect outside the int16 range cannot intersect the 64x64 target
    fb.flags = (uint8_t)((b.blend ? 1 : 0) | (b.rotated ? 2 : 0) | ((b.ax.visible && !off) ? 0 : 4));
    return fb;
}

// Ordered, compacting append of post blits by the whole CTA: candidate k (in the reference's
// submission order) is evaluated by thread k % blockDim; only visible blits are stored, order
// preserved through a warp ballot + a prefix over the warps' counts. make(k, blit, rot) fills the blit.
// Must be called by every thread of the CTA, after a __syncthreads() that follows the frame's initialisation.
template <class F, class MakeFn>
PG2_DEV void emit_post_blits(F& f, int ncand, MakeFn make) {
    const int tid = threadIdx.x, lane = tid % WARP_LANES, warp = tid / WARP_LANES;
    const int nwarps = ((int)blockDim.x + WARP_LANES - 1) / WARP_LANES;
    int n = f.npost;
    for (int base = 0; base < ncand; base += blockDim.x) {
        int k = base + tid;
        Blit b; BlitRot rot;
        b.ax.visible = 0; b.rotated = 0; rot.s = 0.0; rot.c = 1.0;
        if (k < ncand) make(k, b, rot);
        bool vis = k < ncand && b.ax.visible;
        uint32_t m = __ballot_sync(0xffffffffu, vis);
        if (lane == 0) f.wcount[warp] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
        for (int w2 = 0; w2 < nwarps; w2++) { int cnt = f.wcount[w2]; if (w2 < warp) before += cnt; total += cnt; }
        if (vis) {
            int idx = n + before + __popc(m & ((1u << lane) - 1u));
            if (idx < F::MAX_POST) {
                f.fpost[idx] = make_fast(b);
                if (F::ROTATES) f.post_rot[idx] = rot;
            }
        }
        n += total;
        __syncthreads();
    }
    if (tid == 0) f.npost = n < F::MAX_POST ? n : F::MAX_POST;
    __syncthreads();
}

// Tile window of System_Tilemap::render (tilemap.cpp:294-302): inclusive tile index range.
PG2_DEV void tile_window(const Camera& cam, int* lower_x, int* lower_y, int* upper_x, int* upper_y) {
    float hx = __fdiv_rn(__fmul_rn(64.0f, 0.5f), cam.scale);
    float ax = __fmul_rn(__fsub_rn(cam.x, hx), PIXELS_TO_UNIT);
    float ay = __fmul_rn(__fsub_rn(cam.y, hx), PIXELS_TO_UNIT);
    float aw = __fdiv_rn(__fmul_rn(64.0f, PIXELS_TO_UNIT), cam.scale);
    *lower_x = f2i(floorf(ax));
    *lower_y = f2i(floorf(ay));
    *upper_x = f2i(ceilf(__fadd_rn(ax, aw)));
    *upper_y = f2i(ceilf(__fadd_rn(ay, aw)));
}

// ---- per-pixel evaluation ---------------------------------------------------------------------

// After the game's frame builder filled pre/post blits, the tile window, col/row descriptors and tile_tex
// (and __syncthreads()'d): per-column / per-row lookup tables of the tile layer, post-blit bins.
template <class G, class F>
PG2_DEV_NOINLINE void frame_finalize(F& f) {
    const int tid = threadIdx.x;
    for (int i = tid; i < 128 * F::WORDS; i += blockDim.x) (&f.bin[0][0])[i] = 0u;
    for (int i = tid; i < 128; i += blockDim.x) f.bin_any[i] = 0;
    if (tid == 0) { f.wide = 0; f.next_block = 0; }
    // covering ranges: k < 64 -> screen column k, 64..127 -> screen row k-64
    for (int k = tid; k < 128; k += blockDim.x) {
        bool is_row = k >= 64;
        int p = k & 63;
        int n = is_row ? f.nrow : f.ncol;
        int lo = 255, hi = 0;
        for (int cls = 0; cls < f.nclass; cls++) {
            const Axis* ax = is_row ? f.row[cls] : f.col[cls];
            for (int t = 0; t < n; t++) {
                Axis a = ax[t];
                if (a.visible && (unsigned)(p - a.d0) < (unsigned)a.dlen) { lo = min(lo, t); hi = max(hi, t); }
            }
        }
        if (is_row) { f.row_lo[p] = (uint8_t)lo; f.row_hi[p] = (uint8_t)hi; }
        else        { f.col_lo[p] = (uint8_t)lo; f.col_hi[p] = (uint8_t)hi; }
    }
    for (int k = tid; k < f.npre; k += blockDim.x) f.fpre[k] = make_fast(f.pre[k]);
    for (int k = tid; k < f.npre * 128; k += blockDim.x) {
        const Blit& b = f.pre[k >> 7];
        int p = k & 63, is_row = (k >> 6) & 1;
        const Axis& a = is_row ? b.ay : b.ax;
        int32_t v = -1;
        if (b.ax.visible && !b.rotated && (unsigned)(p - a.d0) < (unsigned)a.dlen) {
            int smp = axis_sample(a, p, is_row ? false : (b.flip_h != 0));
            v = is_row ? (int32_t)(b.tex_offset + (uint32_t)smp * b.tex_w) : smp;
        }
        if (is_row) f.pre_row[k >> 7][p] = v; else f.pre_sx[k >> 7][p] = v;
    }
    __syncthreads();
    // per screen column / row and class: source texel of the (at most two) covering tiles
    for (int k = tid; k < 128 * 2 * 2; k += blockDim.x) {
        int p = k & 63, is_row = (k >> 6) & 1, j = (k >> 7) & 1, cls = k >> 8;
        int lo = is_row ? f.row_lo[p] : f.col_lo[p], hi = is_row ? f.row_hi[p] : f.col_hi[p];
        int16_t v = -1;
        if (cls < f.nclass && lo + j <= hi) {
            const Axis& a = is_row ? f.row[cls][lo + j] : f.col[cls][lo + j];
            if (a.visible && (unsigned)(p - a.d0) < (unsigned)a.dlen) v = (int16_t)axis_sample(a, p, false);
        }
        if (is_row) f.row_sy[cls][j][p] = v; else f.col_sx[cls][j][p] = v;
        if (j == 0 && cls == 0 && lo <= hi && hi - lo > 1) f.wide = 1;
    }
    for (int k = tid; k < f.npost; k += blockDim.x) {
        const FastBlit fb = f.fpost[k];
        if (fb.flags & 4u) continue;
        int x0 = fb.x0, y0 = fb.y0, x1 = fb.x0 + fb.w - 1, y1 = fb.y0 + fb.h - 1;
        if (fb.flags & 2u) {   // conservative bounds of a rotated rect: centre +- half diagonal
            int rad = ((int)fb.w + (int)fb.h) / 2 + 2;
            int cx = fb.x0 + fb.w / 2, cy = fb.y0 + fb.h / 2;
            x0 = cx - rad; x1 = cx + rad; y0 = cy - rad; y1 = cy + rad;
        }
        if (x1 < 0 || y1 < 0 || x0 > 63 || y0 > 63) continue;
        x0 = max(x0, 0) >> 3; y0 = max(y0, 0) >> 2; x1 = min(x1, 63) >> 3; y1 = min(y1, 63) >> 2;
        for (int by = y0; by <= y1; by++)
            for (int bx = x0; bx <= x1; bx++) { atomicOr(&f.bin[by * 8 + bx][k >> 5], 1u << (k & 31)); f.bin_any[by * 8 + bx] = 1; }
    }
    if (tid == 0) frame_store_wait();   // the previous frame's bulk store has read f.rgb
    __syncthreads();
}

// Texel of a blit under pixel (X, Y); false when the pixel is not covered.
PG2_DEV bool fast_texel(const FastBlit& fbr, const BlitRot* rot, const uint32_t* __restrict__ atlas, int X, int Y, uint32_t* texel) {
    const FastBlit fb = fbr;
    uint32_t i, j;
    if (fb.flags & 6u) {
        if (fb.flags & 4u) return false;
        // rotated: inverse-map the pixel centre into the un-rotated destination rect (oracle/raster.c step 4)
        double hw = __dmul_rn((double)fb.w, 0.5), hh = __dmul_rn((double)fb.h, 0.5);
        double cx = __dadd_rn((double)fb.x0, hw), cy = __dadd_rn((double)fb.y0, hh);
        double px = __dsub_rn(__dadd_rn((double)X, 0.5), cx);
        double py = __dsub_rn(__dadd_rn((double)Y, 0.5), cy);
        double u = __dadd_rn(__dmul_rn(px, rot->c), __dmul_rn(py, rot->s));
        double v = __dsub_rn(__dmul_rn(py, rot->c), __dmul_rn(px, rot->s));
        double fu = floor(__dadd_rn(u, hw)), fv = floor(__dadd_rn(v, hh));
        if (fu < 0.0 || fv < 0.0 || fu >= (double)fb.w || fv >= (double)fb.h) return false;
        i = (uint32_t)(int)fu; j = (uint32_t)(int)fv;
    } else {
        i = (uint32_t)(X - fb.x0); j = (uint32_t)(Y - fb.y0);
        if (i >= fb.w || j >= fb.h) return false;
    }
    uint32_t sx = (fb.hx + i * fb.incx) >> 16, sy = (fb.hy + j * fb.incy) >> 16;
    *texel = __ldg(atlas + fb.base + sy * fb.tex_w + sx);
    return true;
}

// Effective alpha of a texel of a layer: 255 for opaque-copy textures, else A * alpha_mod / 255.
PG2_DEV uint32_t layer_alpha(uint32_t texel, uint32_t blend, uint32_t alpha_mod) {
    if (!blend) return 255u;
    uint32_t ta = texel >> 24;
    return alpha_mod != 255u ? (ta * alpha_mod) / 255u : ta;
}

// SRC-over of one texel onto a packed 0x00BBGGRR colour; effective alpha 255 replaces and 0 is the identity
// (both exact in blend_texel's integer arithmetic), anything else takes the full per-channel path.
PG2_DEV uint32_t blend_packed(uint32_t color, uint32_t texel, uint32_t blend, uint32_t alpha_mod) {
    uint32_t a = layer_alpha(texel, blend, alpha_mod);
    if (a == 255u) return texel;
    if (a == 0u) return color;
    uint32_t r = color & 255u, g = (color >> 8) & 255u, b = (color >> 16) & 255u;
    blend_texel(r, g, b, texel, blend, alpha_mod);
    return r | g << 8 | b << 16;
}

// clear -> pre -> tiles of one pixel in reference (bottom-up) order; `general` walks the whole covering range
// (frames where more than two tiles cover a column / row), else the two-candidate tables are used.
template <class F>
PG2_DEV_NOINLINE uint32_t shade_base_ordered(const F& f, const uint32_t* __restrict__ atlas, int X, int Y, bool general) {
    uint32_t color = 0u, texel;   // SDL_RenderClear(0,0,0,255)
    for (int k = 0; k < f.npre; k++)
        if (fast_texel(f.fpre[k], nullptr, atlas, X, Y, &texel)) color = blend_packed(color, texel, f.fpre[k].flags & 1u, f.fpre[k].alpha_mod);
    const int rlo = f.row_lo[Y], rhi = f.row_hi[Y], clo = f.col_lo[X], chi = f.col_hi[X];
    for (int ry = rlo; ry <= rhi; ry++)
        for (int cx = clo; cx <= chi; cx++) {
            uint32_t t = f.tile_tex[ry * MAX_WIN + cx];
            if (t == NO_TILE) continue;
            const TileTex tt = f.tiletex[t];
            int sx, sy;
            if (general) {
                const Axis& ax = f.col[tt.cls][cx];
                const Axis& ay = f.row[tt.cls][ry];
                if (!ax.visible || !ay.visible) continue;
                if ((unsigned)(X - ax.d0) >= (unsigned)ax.dlen || (unsigned)(Y - ay.d0) >= (unsigned)ay.dlen) continue;
                sx = axis_sample(ax, X, false); sy = axis_sample(ay, Y, false);
            } else {
                sx = f.col_sx[tt.cls][cx - clo][X]; sy = f.row_sy[tt.cls][ry - rlo][Y];
                if ((sx | sy) < 0) continue;
            }
            texel = __ldg(atlas + tt.offset + (uint32_t)sy * tt.w + (uint32_t)sx);
            color = blend_packed(color, texel, tt.blend, 255u);
        }
    return color;
}

// Shade all 4096 pixels into f.rgb. A warp owns 8x4-pixel blocks (lane = pixel); loop trip counts are uniform over
// the warp (<= 2x2 candidate tiles, the block's post-blit bin) and lanes whose pixel is not covered are predicated
// off, so a warp never serialises different pixels' layer lists.
//   base colour: tile candidates TOP-DOWN ((hi,hi) .. (lo,lo) = reverse painter's order), then the pre blits; the
//     first opaque texel decides the pixel (alpha 255 replaces, alpha 0 is the identity — exact). A partially
//     transparent texel met on the way sends that pixel through shade_base_ordered (reference order) instead.
//   post blits: bottom-up in submission order on top of the base colour.
template <class G, class F>
PG2_DEV_NOINLINE void frame_rasterise(F& f, const uint32_t* __restrict__ atlas) {
    const bool wide = f.wide != 0;
    const int npre = f.npre;
    const int lane = threadIdx.x % WARP_LANES;
    for (;;) {
      // blocks are handed out dynamically: a warp that drew cheap blocks (no sprites) simply takes more of them
      int block = 0;
      if (lane == 0) block = smem_atomic_inc(&f.next_block);
      block = warp_bcast(block);
      if (block >= 128) break;
      for (int l = lane; l < 32; l += WARP_LANES) {
        const int X = ((block & 7) << 3) + (l & 7), Y = ((block >> 3) << 2) + (l >> 3);
        uint32_t color = 0u, texel;
        bool resolved = false, semi = wide;
        if (!wide) {
            const int rlo = f.row_lo[Y], nr = (int)f.row_hi[Y] - rlo + 1;
            const int clo = f.col_lo[X], nc = (int)f.col_hi[X] - clo + 1;
            // every lane walks ITS candidates from the topmost one: candidate index q = jr * ncc + jc, descending
            const int nrr = min(max(nr, 0), 2), ncc = min(max(nc, 0), 2);
            int q = nrr * ncc - 1;
#pragma unroll
            for (int t = 0; t < 4; t++) {
                bool act = !resolved && q >= 0;
                if (!warp_any(act)) break;
                if (act) {
                    const int jr = ncc == 2 ? q >> 1 : q, jc = ncc == 2 ? q & 1 : 0;
                    uint32_t tid = f.tile_tex[(rlo + jr) * MAX_WIN + clo + jc];
                    if (tid != NO_TILE) {
                        const TileTex tt = f.tiletex[tid];
                        int sx = f.col_sx[tt.cls][jc][X], sy = f.row_sy[tt.cls][jr][Y];
                        if ((sx | sy) >= 0) {
                            texel = __ldg(atlas + tt.offset + (uint32_t)sy * tt.w + (uint32_t)sx);
                            uint32_t a = tt.blend ? texel >> 24 : 255u;
                            if (a == 255u) { color = texel; resolved = true; }
                            else if (a != 0u) { semi = true; resolved = true; }
                        }
                    }
                    q--;
                }
            }
            if (warp_any(!resolved)) {
                for (int k = npre - 1; k >= 0; k--) {
                    const int cx = f.pre_sx[k][X], ro = f.pre_row[k][Y];
                    if (!resolved && (cx | ro) >= 0) {
                        texel = __ldg(atlas + (uint32_t)ro + (uint32_t)cx);
                        uint32_t a = layer_alpha(texel, f.fpre[k].flags & 1u, f.fpre[k].alpha_mod);
                        if (a == 255u) { color = texel; resolved = true; }
                        else if (a != 0u) { semi = true; resolved = true; }
                    }
                }
            }
        }
        if (semi) color = shade_base_ordered(f, atlas, X, Y, wide);
        if (f.bin_any[block]) {
            const uint32_t* bins = f.bin[block];
#pragma unroll
            for (int w = 0; w < F::WORDS; w++) {
                uint32_t m = bins[w];
                while (m) {
                    int k = w * 32 + __ffs(m) - 1;
                    m &= m - 1;
                    if (fast_texel(f.fpost[k], &f.post_rot[F::ROTATES ? k : 0], atlas, X, Y, &texel))
                        color = blend_packed(color, texel, f.fpost[k].flags & 1u, f.fpost[k].alpha_mod);
                }
            }
        }
        uint8_t* out = f.rgb + 3 * (Y * OBS_W + X);
        out[0] = (uint8_t)color; out[1] = (uint8_t)(color >> 8); out[2] = (uint8_t)(color >> 16);
      }
    }
}

// The finished frame leaves the SM as ONE 12 288-byte TMA bulk store shared -> global (async proxy). The copy is
// only ISSUED here; the CTA goes on with the next frame's description and waits (frame_store_wait) right before
// it overwrites f.rgb again, so the drain of the staging buffer overlaps useful work.
template <class F>
PG2_DEV void frame_store(F& f, uint8_t* __restrict__ dst) {
#ifdef PG2_HOSTSIM
    memcpy(dst, f.rgb, OBS_BYTES);
    return;
#else
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t src = (uint32_t)__cvta_generic_to_shared(f.rgb);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "n"(OBS_BYTES) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
#endif
}

// Per-CTA table of the game's tile textures (ids < MAX_TILE_TEX), filled once before the first frame.
template <class G, class F>
PG2_DEV void frame_init_tiletex(F& f, const TexInfo* __restrict__ tex) {
    int ntex = 0;
    (void)ntex;
    for (int t = threadIdx.x; t < MAX_TILE_TEX && t < G::NUM_TEX; t += blockDim.x) {
        TexInfo ti = tex[t];
        TileTex tt;
        tt.offset = ti.offset; tt.w = ti.w; tt.blend = ti.blend ? 1 : 0;
        tt.cls = (uint8_t)((G::TILE_CLASSES > 1) ? G::tile_class((uint32_t)t) : 0);
        f.tiletex[t] = tt;
    }
    __syncthreads();
}

}  // namespace pg2
