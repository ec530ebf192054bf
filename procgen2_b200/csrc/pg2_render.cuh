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
constexpr int RENDER_THREADS = 256;
constexpr uint16_t NO_TILE = 0xffff;

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

struct Frame {
    // pre / post blit lists
    Blit pre[MAX_PRE];
    Blit post[MAX_POST];
    BlitRot post_rot[MAX_POST];
    int npre, npost;
    // tile layer: window origin (tile coordinates, y in render space), extents, descriptors per
    // texture shape class (textures of one class share width and height)
    int tx0, ty0, ncol, nrow, nclass;
    Axis col[2][MAX_WIN];
    Axis row[2][MAX_WIN];
    uint16_t tile_tex[MAX_WIN * MAX_WIN];       // texture index per window cell or NO_TILE
    uint8_t col_lo[OBS_W], col_hi[OBS_W];       // per screen column: range of tile columns covering it
    uint8_t row_lo[OBS_H], row_hi[OBS_H];       // (lo > hi: none)
    // post-blit binning: 8x8-pixel blocks x 128 blits
    uint32_t bin[64][MAX_POST / 32];
    // staged output frame
    alignas(16) uint8_t rgb[OBS_BYTES];
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

// Ordered, compacting append of post blits by the first warp of the CTA: candidate k (in the
// reference's submission order) is evaluated by lane k % 32; only visible blits are stored,
// order preserved through a ballot prefix. make(k, blit, rot) fills the blit.
template <class MakeFn>
PG2_DEV void emit_post_blits(Frame& f, int ncand, MakeFn make) {
    if ((int)threadIdx.x >= WARP_LANES) return;
    const int lane = threadIdx.x;
    int n = f.npost;
    for (int base = 0; base < ncand; base += WARP_LANES) {
        int k = base + lane;
        Blit b; BlitRot rot;
        b.ax.visible = 0; b.rotated = 0; rot.s = 0.0; rot.c = 1.0;
        if (k < ncand) make(k, b, rot);
        bool vis = k < ncand && b.ax.visible;
        uint32_t m = __ballot_sync(0xffffffffu, vis);
        if (vis) {
            int idx = n + __popc(m & ((1u << lane) - 1u));
            if (idx < MAX_POST) { f.post[idx] = b; f.post_rot[idx] = rot; }
        }
        n += __popc(m);
    }
    if (lane == 0) f.npost = n < MAX_POST ? n : MAX_POST;
    __syncwarp();
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

PG2_DEV void shade_blit(const Blit& b, const BlitRot* rot, const uint32_t* __restrict__ atlas,
                                           int X, int Y, uint32_t& r, uint32_t& g, uint32_t& bl) {
    int sx, sy;
    if (!b.rotated) {
        if ((unsigned)(X - b.ax.d0) >= (unsigned)b.ax.dlen || (unsigned)(Y - b.ay.d0) >= (unsigned)b.ay.dlen) return;
        sx = axis_sample(b.ax, X, b.flip_h);
        sy = axis_sample(b.ay, Y, false);
    } else {
        // inverse-map the pixel centre into the un-rotated destination rect (oracle/raster.c step 4)
        double hw = __dmul_rn((double)b.ax.dlen, 0.5), hh = __dmul_rn((double)b.ay.dlen, 0.5);
        double cx = __dadd_rn((double)b.ax.d0, hw), cy = __dadd_rn((double)b.ay.d0, hh);
        double px = __dsub_rn(__dadd_rn((double)X, 0.5), cx);
        double py = __dsub_rn(__dadd_rn((double)Y, 0.5), cy);
        double u = __dadd_rn(__dmul_rn(px, rot->c), __dmul_rn(py, rot->s));
        double v = __dsub_rn(__dmul_rn(py, rot->c), __dmul_rn(px, rot->s));
        double fu = floor(__dadd_rn(u, hw)), fv = floor(__dadd_rn(v, hh));
        if (fu < 0.0 || fv < 0.0 || fu >= (double)b.ax.dlen || fv >= (double)b.ay.dlen) return;
        int i = (int)fu, j = (int)fv;
        sx = b.ax.s0 + (int)((b.ax.inc / 2u + (uint32_t)i * b.ax.inc) >> 16);
        sy = b.ay.s0 + (int)((b.ay.inc / 2u + (uint32_t)j * b.ay.inc) >> 16);
    }
    uint32_t texel = __ldg(atlas + b.tex_offset + (uint32_t)sy * b.tex_w + (uint32_t)sx);
    blend_texel(r, g, bl, texel, b.blend, b.alpha_mod);
}

// Conservative screen-space bounds of a blit (exact for axis-aligned ones).
PG2_DEV void blit_bounds(const Blit& b, int* x0, int* y0, int* x1, int* y1) {
    if (!b.rotated) {
        *x0 = b.ax.d0; *x1 = b.ax.d0 + b.ax.dlen - 1; *y0 = b.ay.d0; *y1 = b.ay.d0 + b.ay.dlen - 1;
    } else {
        int rad = (b.ax.dlen + b.ay.dlen) / 2 + 2;   // >= half diagonal
        int cx = b.ax.d0 + b.ax.dlen / 2, cy = b.ay.d0 + b.ay.dlen / 2;
        *x0 = cx - rad; *x1 = cx + rad; *y0 = cy - rad; *y1 = cy + rad;
    }
}

// After the game's frame builder filled pre/post blits, the tile window, col/row descriptors
// and tile_tex (and __syncthreads()'d), derive the per-column / per-row candidate lists and
// the post-blit bins.
PG2_DEV_NOINLINE void frame_finalize(Frame& f) {
    int tid = threadIdx.x;
    for (int i = tid; i < 64 * (MAX_POST / 32); i += blockDim.x) (&f.bin[0][0])[i] = 0u;
    // covering ranges: thread t < 64 -> screen column t, 64..127 -> screen row t-64
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
    __syncthreads();
    for (int k = tid; k < f.npost; k += blockDim.x) {
        const Blit& b = f.post[k];
        if (!b.ax.visible) continue;
        int x0, y0, x1, y1;
        blit_bounds(b, &x0, &y0, &x1, &y1);
        x0 = max(x0, 0) >> 3; y0 = max(y0, 0) >> 3; x1 = min(x1, 63) >> 3; y1 = min(y1, 63) >> 3;
        for (int by = y0; by <= y1; by++)
            for (int bx = x0; bx <= x1; bx++) atomicOr(&f.bin[by * 8 + bx][k >> 5], 1u << (k & 31));
    }
    __syncthreads();
}

// Shade all 4096 pixels into f.rgb. `texinfo` = the game's texture table, `tile_class[tex]`
// is implied by TexInfo shape through the game's CLASS_OF callback (template parameter).
template <class G>
PG2_DEV_NOINLINE void frame_rasterise(Frame& f, const TexInfo* __restrict__ texinfo, const uint32_t* __restrict__ atlas) {
    for (int p = threadIdx.x; p < OBS_W * OBS_H; p += blockDim.x) {
        int X = p & 63, Y = p >> 6;
        uint32_t r = 0, g = 0, b = 0;   // SDL_RenderClear(0,0,0,255)
        for (int k = 0; k < f.npre; k++)
            if (f.pre[k].ax.visible) shade_blit(f.pre[k], nullptr, atlas, X, Y, r, g, b);
        // tile layer: y-major, x-minor painter's order
        {
            int rlo = f.row_lo[Y], rhi = f.row_hi[Y], clo = f.col_lo[X], chi = f.col_hi[X];
            for (int ry = rlo; ry <= rhi; ry++)
                for (int cx = clo; cx <= chi; cx++) {
                    uint32_t tex = f.tile_tex[ry * MAX_WIN + cx];
                    if (tex == NO_TILE) continue;
                    int cls = (G::TILE_CLASSES > 1) ? G::tile_class(tex) : 0;
                    const Axis& ax = f.col[cls][cx];
                    const Axis& ay = f.row[cls][ry];
                    if (!ax.visible || !ay.visible) continue;
                    if ((unsigned)(X - ax.d0) >= (unsigned)ax.dlen || (unsigned)(Y - ay.d0) >= (unsigned)ay.dlen) continue;
                    TexInfo ti = texinfo[tex];
                    int sx = axis_sample(ax, X, false), sy = axis_sample(ay, Y, false);
                    uint32_t texel = __ldg(atlas + ti.offset + (uint32_t)sy * ti.w + (uint32_t)sx);
                    blend_texel(r, g, b, texel, ti.blend, 255u);
                }
        }
        // post blits through the 8x8 bins, ascending index = submission order
        const uint32_t* bins = f.bin[(Y >> 3) * 8 + (X >> 3)];
#pragma unroll
        for (int w = 0; w < MAX_POST / 32; w++) {
            uint32_t m = bins[w];
            while (m) {
                int k = w * 32 + __ffs(m) - 1;
                m &= m - 1;
                shade_blit(f.post[k], &f.post_rot[k], atlas, X, Y, r, g, b);
            }
        }
        f.rgb[3 * p + 0] = (uint8_t)r;
        f.rgb[3 * p + 1] = (uint8_t)g;
        f.rgb[3 * p + 2] = (uint8_t)b;
    }
}

// One 12 288-byte TMA bulk store shared -> global (Blackwell/Hopper async proxy).
PG2_DEV void frame_store(Frame& f, uint8_t* __restrict__ dst) {
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
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem may be reused afterwards
    }
    __syncthreads();
#endif
}

}  // namespace pg2
