// Observation rasteriser: one CTA renders one environment's 64x64x3 frame.
//
// Replaces render_game() (games/coinrun/coinrun.cpp:443-470 and its six siblings) together with
// the SDL3 software blits it issues (SDL_RenderTextureRotated, renderer.cpp:78/97) and the
// RGBA->RGB pack loop (coinrun.cpp:377-388). Draw order is the reference's painter's order:
//   clear(0,0,0) -> "pre" blit (background) -> tile layer (y-major, x-minor; tilemap.cpp:303-320)
//   -> "post" blits (particles, sprites, agent, HUD) in submission order.
//
// The frame is drawn in shared memory by row bands (8 bands x 8 rows, handed out to the CTA's warps):
//   base pass   every thread owns a run of 4 pixels of one row. Background + tile layer are a GATHER: per pixel
//               the top-most tile candidate is found from per-row tile presence bitmaps (no walk over empty
//               cells), ONE texel is fetched (four independent fetches in flight per thread), an opaque texel
//               decides the pixel; a transparent / translucent one sends that pixel to the ordered slow path.
//               Four finished pixels are packed into three 32-bit words with byte permutes and stored.
//   post pass   the same warp then draws the post blits that touch its band, blit by blit in submission order
//               (lanes = an 8x4 patch of the blit's destination rectangle), straight onto the packed RGB rows.
//   store       the band leaves the SM as one 1 536-byte TMA bulk store (cp.async.bulk.global.shared::cta).
//
// The tile layer exploits that render_texture() is separable: a tile's destination columns only
// depend on its x index and its rows only on its y index, so a frame needs <= 32 column and <= 32
// row descriptors instead of up to 27x27 blit records; per screen column / row the (at most two) covering
// tile columns / rows and their source texel indices are tabulated once per frame (ColDesc / RowDesc).
#pragma once
#include "pg2_common.cuh"

namespace pg2 {

constexpr int MAX_WIN = 32;        // tile window extent per axis (maze: 27)
constexpr int MAX_PRE = 2;
#ifndef PG2_RENDER_THREADS
#define PG2_RENDER_THREADS 128
#endif
constexpr int RENDER_THREADS = PG2_RENDER_THREADS;
constexpr uint8_t NO_TILE = 0xff;
constexpr int BAND_ROWS = 8, NUM_BANDS = OBS_H / BAND_ROWS, BAND_BYTES = BAND_ROWS * OBS_W * 3;

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

struct BlitRot {
    double s, c;       // sin/cos of the blit angle (deterministic, see sincos_deg)
    int ex, ey;        // half extents (pixels, rounded up + 2) of the rotated rect's axis-aligned bounding box
};

// Per-pixel form of a blit (what the frame keeps): coverage test = two unsigned compares, sampling = one
// multiply-add + shift per axis (a horizontal flip is folded into hx / incx, which wrap modulo 2^32 to the
// exact non-negative value), texel address = base + sy * tex_w + sx.
struct alignas(16) FastBlit {
    int16_t x0, y0; uint16_t w, h;              // integer destination rect (SDL truncates the float rect)
    uint32_t hx, incx;                          // sx = (hx + i * incx) >> 16
    uint32_t hy, incy, base;                    // base = tex_offset + s0y * tex_w + s0x
    uint16_t tex_w; uint8_t flags, alpha_mod;   // flags: 1 blend, 2 rotated, 4 invisible
};

constexpr int MAX_TILE_TEX = 32;

// Tile layer + background under one screen column: clo = first covering tile column of the window; per texture
// shape class (textures of one class share width and height) and candidate j (tile column clo + j) a validity bit
// and the source texel x. vc masks are in "candidate index" form (candidate q = jr * 2 + jc): 0b0101 for jc = 0,
// 0b1010 for jc = 1, so that (cw & rw) >> (8 + 4 * cls) is the set of candidates whose class-cls axes cover the pixel.
// Stored as three word arrays indexed by col_slot(X), so that the 16 lanes of a row (lane l owns columns 4l .. 4l+3)
// read consecutive words: no bank conflicts.
struct ColDesc {
    uint32_t cw;       // clo | vc[0] << 8 | vc[1] << 12
    uint32_t csx;      // sx[cls][j] in byte cls * 2 + j
    int32_t pre_sx;    // background: source x under this column, -1: not covered
};
PG2_DEV int col_slot(int X) { return (X >> 2) | (X & 3) << 4; }
// Window cell: atlas offset of the tile's texture | blend << 28 | shape class << 29 | 1 << 31; 0 = no tile.
constexpr uint32_t CELL_OFFSET_MASK = 0x0fffffffu, CELL_PRESENT = 0x80000000u;
// Same for one screen row, plus the tile presence bitmaps of the two candidate tile rows (bit = tile column).
struct alignas(16) RowDesc {
    uint32_t rw;       // rlo | vr[0] << 8 | vr[1] << 12 | rlo * MAX_WIN << 16   (vr: 0b0011 for jr = 0, 0b1100 for jr = 1)
    int32_t pre_row;   // background: tex_offset + sy * tex_w under this row, -1: not covered
    uint32_t syw[2];   // [cls]: (sy * tex_w) of candidate row 0 | candidate row 1 << 16
    uint32_t rm[2][2]; // [cls][jr]: presence bitmap of class-cls tiles in tile row rlo + jr
};

// Frame description of ONE environment, in shared memory. MAXP = capacity of the post-blit list (per game),
// ROT = whether the game ever rotates a blit (bossfight, caveflyer, jumper HUD).
template <int MAXP, bool ROT, int NCLS>
struct FrameT {
    static constexpr int MAX_POST = MAXP, NROT = ROT ? MAXP : 1;
    static constexpr bool ROTATES = ROT;
    alignas(16) uint8_t band_rgb[RENDER_THREADS / 32][BAND_BYTES];   // per warp: the band it is drawing, packed RGB rows
    // ---- the view: everything the rasteriser needs of background + tile layer
    alignas(16) uint32_t col_cw[OBS_W];         // ColDesc fields, indexed by col_slot(X)
    uint32_t col_csx[OBS_W];
    int32_t col_pre[OBS_W];
    RowDesc rowd[OBS_H];
    uint32_t cell[(MAX_WIN + 1) * MAX_WIN];     // window cells (+1 row: branch-free reads)
    FastBlit fpre[MAX_PRE];
    int npre;
    int wide;                                   // the frame needs the general ordered path for every pixel (never observed)
    int pre_blend;                              // background texture carries alpha
    // ---- per frame
    int reuse;                                  // the base image comes from the env's cache: the view is not built
    FastBlit fpost[MAXP];
    BlitRot post_rot[NROT];
    Blit pre[MAX_PRE];
    int npost;
    // tile layer: window origin (tile coordinates, y in render space), extents, descriptors per texture shape class
    int tx0, ty0, ncol, nrow, nclass;
    Axis col[NCLS][MAX_WIN];
    Axis row[NCLS][MAX_WIN];
    uint32_t rowmask[2][MAX_WIN + 1];           // [cls][tile row]: bit cx set = a class-cls tile at window column cx
    int cov_lo[2 * OBS_W], cov_hi[2 * OBS_W];   // [0,64): per screen column, [64,128): per screen row: covering tile range
    uint16_t bandmask[MAXP];                    // post blit k touches band b <=> bit b
    uint8_t live[256];                          // live_list(): ids of the live sprites in set order
    int wcount[2][RENDER_THREADS / 32];         // emit_post_blits: visible blits per warp (double-buffered by round)
    int next_band;                              // dynamic hand-out of the row bands to warps
    int class_w[2];                             // texture width of the tile shape classes
    uint32_t tileword[MAX_TILE_TEX];            // per CTA (filled once): window cell word of every tile texture id
};

// Deterministic sin/cos in degrees, mirrored operation by operation from oracle/raster.c
// (pg2o_sincos_deg): IEEE double add/mul only, fixed order, no FMA.
PG2_DEV_CALL void sincos_deg(double deg, double* s, double* c) {
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
    make_axis_xy(px, py, cam.x, cam.y, cam.scale, t.w, t.h, scale, flip_h, &b.ax, &b.ay);
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

// Half extents of the axis-aligned bounding box of a rotated destination rect: a pixel centre passes the inverse-mapping
// test of rotated_texel_coords only if |x - cx| <= hw |c| + hh |s| and |y - cy| <= hw |s| + hh |c|; + 2 px of slack for
// the integer centre and rounding (a bound only has to be conservative, it never changes which pixels are drawn).
PG2_DEV void rotated_extents(const FastBlit& fb, BlitRot* rot) {
    const double hw = 0.5 * (double)fb.w, hh = 0.5 * (double)fb.h, ac = fabs(rot->c), as = fabs(rot->s);
    rot->ex = (int)ceil(hw * ac + hh * as) + 2;
    rot->ey = (int)ceil(hw * as + hh * ac) + 2;
}

// Bands (8 rows each) a blit can touch.
PG2_DEV uint32_t blit_bands(const FastBlit& fb, const BlitRot& rot) {
    int y0 = fb.y0, y1 = fb.y0 + fb.h - 1;
    if (fb.flags & 2u) {
        const int cy = fb.y0 + fb.h / 2;
        y0 = cy - rot.ey; y1 = cy + rot.ey;
    }
    if (y1 < 0 || y0 >= OBS_H) return 0u;
    int b0 = max(y0, 0) / BAND_ROWS, b1 = min(y1, OBS_H - 1) / BAND_ROWS;
    return ((2u << b1) - 1u) & ~((1u << b0) - 1u);
}

// What a game's frame builder asks for per post-blit candidate. The expensive part (make_blit and friends) is run by
// emit_post_blits for all lanes of a warp TOGETHER, whatever kind of sprite each lane describes.
struct BlitReq {
    int mode;            // 0: nothing to draw, 1: render_texture, 2: render_texture_rotated, 3: explicit rect + angle
    int tex_id;
    float x, y, w, h;    // mode 1/2: world position in pixels (w, h unused); mode 3: float destination rect
    float scale, alpha, rotation;
    double angle_deg;
    bool flip;
    Camera cam;
    PG2_DEV void plain(int t, float px, float py, const Camera& c, float sc, float al = 1.0f, bool fl = false) {
        mode = 1; tex_id = t; x = px; y = py; cam = c; scale = sc; alpha = al; flip = fl;
    }
    PG2_DEV void rotated(int t, float px, float py, const Camera& c, float rot, float sc, float al, BlitRot*) {
        mode = 2; tex_id = t; x = px; y = py; cam = c; rotation = rot; scale = sc; alpha = al; flip = false;
    }
    PG2_DEV void rect(int t, float dx, float dy, float dw, float dh, double ang, BlitRot*) {
        mode = 3; tex_id = t; x = dx; y = dy; w = dw; h = dh; angle_deg = ang; flip = false;
    }
};

// Ordered, compacting append of post blits by the whole CTA: candidate k (in the reference's
// submission order) is described by thread k % blockDim (make(k, req, rot) fills the request or leaves it empty);
// only visible blits are stored, order preserved through a warp ballot + a prefix over the warps' counts.
// Must be called by every thread of the CTA, once per frame; ONE barrier per 128 candidates.
template <class F, class MakeFn>
PG2_DEV void emit_post_blits(F& f, const TexInfo* tex, int ncand, MakeFn make) {
    const int tid = threadIdx.x, lane = tid % WARP_LANES, warp = tid / WARP_LANES;
    const int nwarps = ((int)blockDim.x + WARP_LANES - 1) / WARP_LANES;
    int n = 0, round = 0;
    for (int base = 0; base < ncand; base += blockDim.x, round ^= 1) {
        int k = base + tid;
        BlitReq req; BlitRot rot;
        req.mode = 0; rot.s = 0.0; rot.c = 1.0; rot.ex = 0; rot.ey = 0;
        if (k < ncand) make(k, req, rot);
        Blit b;
        b.ax.visible = 0;
        if (req.mode == 1) b = make_blit(tex, req.tex_id, req.x, req.y, req.cam, req.scale, req.alpha, req.flip);
        else if (req.mode == 2) b = make_blit_rotated(tex, req.tex_id, req.x, req.y, req.cam, req.rotation, req.scale, req.alpha, &rot);
        else if (req.mode == 3) b = make_blit_rect(tex, req.tex_id, req.x, req.y, req.w, req.h, req.angle_deg, &rot);
        bool vis = req.mode != 0 && b.ax.visible;
        FastBlit fb;
        if (vis) { fb = make_fast(b); vis = !(fb.flags & 4u); }
        uint32_t m = __ballot_sync(0xffffffffu, vis);
        if (lane == 0) f.wcount[round][warp] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
        for (int w2 = 0; w2 < nwarps; w2++) { int cnt = f.wcount[round][w2]; if (w2 < warp) before += cnt; total += cnt; }
        if (vis) {
            int idx = n + before + __popc(m & ((1u << lane) - 1u));
            if (idx < F::MAX_POST) {
                if (F::ROTATES && (fb.flags & 2u)) rotated_extents(fb, &rot);
                f.fpost[idx] = fb;
                f.bandmask[idx] = (uint8_t)blit_bands(fb, rot);
                if (F::ROTATES) f.post_rot[idx] = rot;
            }
        }
        n += total;
    }
    if (tid == 0) f.npost = n < F::MAX_POST ? n : F::MAX_POST;
}

// f.live[0 .. count) = the ids id_or_neg(j) >= 0, j < n, in order (the live members of an ECS set whose dead entries
// are skipped by the reference's iteration). Every warp of the CTA computes the whole list redundantly — identical
// values to identical addresses — so no CTA barrier is needed: a warp reads the list after its own __syncwarp().
template <class F, class Fn>
PG2_DEV int live_list(F& f, int n, Fn id_or_neg) {
    const int lane = threadIdx.x % WARP_LANES;
    int cnt = 0;
    for (int base = 0; base < n; base += WARP_LANES) {
        const int j = base + lane;
        const int v = j < n ? id_or_neg(j) : -1;
        const uint32_t m = lane_ballot(v >= 0, lane);
        if (v >= 0 && cnt < 256) f.live[(cnt + __popc(m & ((1u << lane) - 1u))) & 255] = (uint8_t)v;
        cnt += __popc(m);
    }
    __syncwarp();
    return cnt;
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

// Games whose camera and tile map are fixed within an episode (G::STATIC_VIEW: maze, chaser) keep the BASE IMAGE of an env
// (clear + background + tile layer, 12 288 B of packed RGB) in HBM: the first frame of an episode draws and stores it, the
// following frames load it band by band instead of describing and rasterising the tile layer again.
constexpr size_t VIEW_CACHE_BYTES = OBS_BYTES;

// Start of a frame (every thread; followed by a __syncthreads() before the game's frame builder runs). `reuse`: the base
// image comes from the env's cache (G::STATIC_VIEW), so the view is not built.
template <class F>
PG2_DEV void frame_begin(F& f, bool reuse = false) {   // `reuse` is only looked at by thread 0 (which knows the env)
    const int tid = threadIdx.x;
    if (tid == 0) {
        f.npost = 0; f.ncol = 0; f.nrow = 0; f.nclass = 1; f.next_band = 0; f.reuse = reuse ? 1 : 0;
        if (!reuse) { f.npre = 0; f.wide = 0; f.pre_blend = 0; }
    }
    for (int k = tid; k < 2 * OBS_W; k += blockDim.x) { f.cov_lo[k] = 255; f.cov_hi[k] = -1; }
}

// One band of the base image: band buffer <-> cache (16-byte words, 3 per lane, coalesced).
struct alignas(16) Word16 { uint32_t a, b, c, d; };
PG2_DEV void band_from_cache(uint8_t* buf, const uint8_t* __restrict__ cache_band, int lane) {
    for (int i = lane; i < BAND_BYTES / 16; i += WARP_LANES) ((Word16*)buf)[i] = ((const Word16*)cache_band)[i];
}
PG2_DEV void band_to_cache(const uint8_t* buf, uint8_t* __restrict__ cache_band, int lane) {
    for (int i = lane; i < BAND_BYTES / 16; i += WARP_LANES) ((Word16*)cache_band)[i] = ((const Word16*)buf)[i];
}

// Background + tile layer of a frame (the "pre" blit of render_game and System_Tilemap::render, tilemap.cpp:294-320),
// called by every thread of the CTA from the game's frame builder:
//   - the two axes of the background blit (texture bg_tex at world pixel position (bg_x, bg_y), scale bg_scale) and the
//     window's column / row axes per texture shape class (class_tex(cls) = a texture of that class) are independent
//     make_axis jobs: one job per thread, counted down from the CTA's LAST thread (the first warps build post blits
//     meanwhile), all running the same instruction stream;
//   - the window's cells (tile_at(x, y): tile texture id or NO_TILE; x = window column + lx, y = render-space tile row)
//     and the per-row presence bitmaps, one warp per tile row.
// Every tile axis also registers itself in the covering range of the screen columns / rows it touches
// (cov_lo / cov_hi, initialised by frame_begin).
template <class F, class ClassTex, class TileAt>
PG2_DEV void build_tile_layer(F& f, const Camera& cam, const TexInfo* tex, int nclass, int lx, int ly, int ncol, int nrow,
                              ClassTex class_tex, TileAt tile_at, int bg_tex, float bg_x, float bg_y, float bg_scale) {
    const int tid = threadIdx.x, lane = tid % WARP_LANES, warp = tid / WARP_LANES;
    const int nwarps = ((int)blockDim.x + WARP_LANES - 1) / WARP_LANES;
    if (f.reuse) return;   // the base image comes from the env's cache: no descriptors, cells or background needed
    const int per = ncol + nrow, njobs = 2 + nclass * per;
    if (tid == 0) { f.tx0 = lx; f.ty0 = ly; f.ncol = ncol; f.nrow = nrow; f.nclass = nclass; f.npre = 1; }
    for (int job = (int)blockDim.x - 1 - tid; job < njobs; job += blockDim.x) {
        const bool bg = job < 2;
        const int t = job - 2, cls = (!bg && t >= per) ? 1 : 0, u = bg ? 0 : t - cls * per;
        const bool is_row = bg ? job == 1 : u >= ncol;
        const int idx = is_row ? u - ncol : u;
        const TexInfo ti = tex[bg ? bg_tex : class_tex(cls)];
        const float scale = bg ? bg_scale : __fdiv_rn(UNIT_TO_PIXELS, (float)ti.w);
        const float pos = bg ? (is_row ? bg_y : bg_x) : __fmul_rn((float)((is_row ? ly : lx) + idx), UNIT_TO_PIXELS);
        const Axis a = make_axis(pos, is_row ? cam.y : cam.x, cam.scale, 64.0f, is_row ? ti.h : ti.w, scale, false, is_row);
        if (bg) {
            if (is_row) f.pre[0].ay = a;
            else {   // the x-axis job also fills the rest of the blit (make_blit, alpha 1, no flip)
                f.pre[0].ax = a;
                f.pre[0].tex_offset = ti.offset; f.pre[0].tex_w = ti.w; f.pre[0].blend = (uint8_t)ti.blend;
                f.pre[0].alpha_mod = 255; f.pre[0].flip_h = 0; f.pre[0].rotated = 0;
            }
            continue;
        }
        if (is_row) f.row[cls][idx] = a; else f.col[cls][idx] = a;
        if (a.visible && a.d0 > -65536 && a.d0 < 65536 && a.dlen < 65536) {   // register in the covering ranges
            const int p0 = max(a.d0, 0), p1 = min(a.d0 + a.dlen - 1, OBS_W - 1), o = is_row ? OBS_W : 0;
            for (int pp = p0; pp <= p1; pp++) { atomicMin(&f.cov_lo[o + pp], idx); atomicMax(&f.cov_hi[o + pp], idx); }
        }
    }
    for (int cls = tid; cls < nclass; cls += blockDim.x) f.class_w[cls] = tex[class_tex(cls)].w;
    // window cells: one warp per tile row, lanes = tile columns (rows >= nrow are never referenced: the always-empty
    // row MAX_WIN stands in for them, see frame_finalize)
    for (int ry = warp; ry < nrow; ry += nwarps) {
        uint32_t m0 = 0u, m1 = 0u;
        for (int cx = lane; cx < MAX_WIN; cx += WARP_LANES) {
            const uint32_t tt = cx < ncol ? (uint32_t)tile_at(lx + cx, ly + ry) : (uint32_t)NO_TILE;
            const uint32_t w = tt != NO_TILE ? f.tileword[tt & (MAX_TILE_TEX - 1)] : 0u;
            f.cell[ry * MAX_WIN + cx] = w;
            m0 |= lane_ballot(w != 0u && !(w >> 29 & 1u), cx);
            m1 |= lane_ballot((w >> 29 & 1u) != 0u, cx);
        }
        if (lane == 0) { f.rowmask[0][ry] = m0; f.rowmask[1][ry] = m1; }
    }
}

// ---- per-pixel evaluation ---------------------------------------------------------------------

// After the game's frame builder (and a __syncthreads()): ColDesc / RowDesc of every screen column / row.
template <class G, class F>
PG2_DEV_NOINLINE void frame_finalize(F& f) {
    if (f.reuse) { __syncthreads(); return; }
    const int tid = threadIdx.x;
    const int npre = f.npre, nclass = f.nclass;
    // the fast path handles ONE un-rotated background with alpha_mod 255
    const bool pre_ok = npre == 0 || (npre == 1 && !f.pre[0].rotated && f.pre[0].alpha_mod == 255);
    if (tid == 0) {
        if (!pre_ok) f.wide = 1;
        f.pre_blend = npre >= 1 ? f.pre[npre - 1].blend : 0;
    }
    for (int k = tid; k < npre; k += blockDim.x) {
        Blit b = f.pre[k];
        if (!b.ay.visible) b.ax.visible = 0;   // make_blit's rule (the two axes were built by different threads)
        f.fpre[k] = make_fast(b);
    }
    for (int k = tid; k < 2 * OBS_W; k += blockDim.x) {
        const bool is_row = k >= OBS_W;
        const int p = k & (OBS_W - 1);
        const int lo = f.cov_lo[k], hi = f.cov_hi[k];
        uint32_t word = 0u, smp[2] = { 0u, 0u };
        if (lo <= hi) {
            word = (uint32_t)lo;
            if (hi - lo > 1) f.wide = 1;
            for (int cls = 0; cls < nclass; cls++)
                for (int j = 0; j < 2; j++) {
                    if (lo + j > hi) continue;
                    const Axis& a = is_row ? f.row[cls][lo + j] : f.col[cls][lo + j];
                    if (!a.visible || (unsigned)(p - a.d0) >= (unsigned)a.dlen) continue;
                    const uint32_t v = (uint32_t)axis_sample(a, p, false);
                    if (is_row) {
                        const uint32_t w = (uint32_t)f.class_w[cls];
                        if (v * w > 0xffffu) f.wide = 1;
                        smp[cls] |= ((v * w) & 0xffffu) << (16 * j);
                        word |= (j ? 0xcu : 0x3u) << (8 + 4 * cls);
                    } else {
                        if (v > 0xffu) f.wide = 1;
                        smp[0] |= (v & 0xffu) << (8 * (cls * 2 + j));
                        word |= (j ? 0xau : 0x5u) << (8 + 4 * cls);
                    }
                }
        }
        int32_t pv = -1;
        if (npre >= 1 && pre_ok) {
            const Blit& b = f.pre[0];
            const Axis& a = is_row ? b.ay : b.ax;
            if (b.ax.visible && b.ay.visible && (unsigned)(p - a.d0) < (unsigned)a.dlen) {
                int s = axis_sample(a, p, is_row ? false : (b.flip_h != 0));
                pv = is_row ? (int32_t)(b.tex_offset + (uint32_t)s * b.tex_w) : s;
            }
        }
        if (is_row) {
            RowDesc rd;
            rd.rw = word | ((lo <= hi ? (uint32_t)lo : 0u) * MAX_WIN) << 16; rd.pre_row = pv; rd.syw[0] = smp[0]; rd.syw[1] = smp[1];
            const int r0 = lo <= hi ? lo : MAX_WIN;       // row MAX_WIN is always empty
            const int r1 = lo < hi ? lo + 1 : MAX_WIN;
            rd.rm[0][0] = f.rowmask[0][r0]; rd.rm[0][1] = f.rowmask[0][r1];
            rd.rm[1][0] = f.rowmask[1][r0]; rd.rm[1][1] = f.rowmask[1][r1];
            f.rowd[p] = rd;
        } else {
            const int slot = col_slot(p);
            f.col_cw[slot] = word; f.col_csx[slot] = smp[0]; f.col_pre[slot] = pv;
        }
    }
    __syncthreads();
}

// Rotated blit: inverse-map the pixel centre into the un-rotated destination rect (oracle/raster.c step 4).
// (Measured: a real call here costs bossfight / caveflyer ~10 % of the render.)
PG2_DEV bool rotated_texel_coords(int x0, int y0, int w, int h, double rs, double rc, int X, int Y, uint32_t* oi, uint32_t* oj) {
    double hw = __dmul_rn((double)w, 0.5), hh = __dmul_rn((double)h, 0.5);
    double cx = __dadd_rn((double)x0, hw), cy = __dadd_rn((double)y0, hh);
    double px = __dsub_rn(__dadd_rn((double)X, 0.5), cx);
    double py = __dsub_rn(__dadd_rn((double)Y, 0.5), cy);
    double u = __dadd_rn(__dmul_rn(px, rc), __dmul_rn(py, rs));
    double v = __dsub_rn(__dmul_rn(py, rc), __dmul_rn(px, rs));
    double fu = floor(__dadd_rn(u, hw)), fv = floor(__dadd_rn(v, hh));
    if (fu < 0.0 || fv < 0.0 || fu >= (double)w || fv >= (double)h) return false;
    *oi = (uint32_t)(int)fu; *oj = (uint32_t)(int)fv;
    return true;
}

// Texel of a blit under pixel (X, Y); false when the pixel is not covered.
template <bool ROT>
PG2_DEV bool fast_texel(const FastBlit& fb, const BlitRot* rot, const uint32_t* __restrict__ atlas, int X, int Y, uint32_t* texel) {
    uint32_t i, j;
    if (ROT && (fb.flags & 2u)) {
        if (!rotated_texel_coords(fb.x0, fb.y0, fb.w, fb.h, rot->s, rot->c, X, Y, &i, &j)) return false;
    } else {
        i = (uint32_t)(X - fb.x0); j = (uint32_t)(Y - fb.y0);
        if (i >= fb.w || j >= fb.h) return false;
    }
    uint32_t sx = (fb.hx + i * fb.incx) >> 16, sy = (fb.hy + j * fb.incy) >> 16;
    *texel = __ldg(atlas + fb.base + sy * fb.tex_w + sx);
    return true;
}

// Effective alpha of a texel of a layer: A * alpha_mod / 255. Opaque-copy (RGB) textures are stored with A = 255 in
// the atlas (assets.cpp), and SRC-over with alpha 255 IS the copy, so the blend flag never needs to be consulted.
PG2_DEV uint32_t layer_alpha(uint32_t texel, uint32_t /*blend*/, uint32_t alpha_mod) {
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

// Tile candidates of a pixel as a 4-bit set (candidate q = jr * 2 + jc <=> tile (rlo + jr, clo + jc)): a tile is
// there and the axes of its shape class cover the pixel. Painter's order = ascending q.
template <int NCLASS>
PG2_DEV uint32_t tile_candidates(const RowDesc& rd, uint32_t cw) {
    const uint32_t clo = cw & 31u, v = cw & rd.rw;
    uint32_t p = (((rd.rm[0][0] >> clo) & 3u) | ((rd.rm[0][1] >> clo) & 3u) << 2) & (v >> 8);
    if (NCLASS > 1) p |= (((rd.rm[1][0] >> clo) & 3u) | ((rd.rm[1][1] >> clo) & 3u) << 2) & (v >> 12);
    return p & 15u;
}

template <class F>
PG2_DEV ColDesc load_col(const F& f, int X) {
    const int slot = col_slot(X);
    ColDesc cd;
    cd.cw = f.col_cw[slot]; cd.csx = f.col_csx[slot]; cd.pre_sx = f.col_pre[slot];
    return cd;
}

// Atlas index of tile candidate q under a pixel. q = 2 jr + jc, so the cell (rlo + jr, clo + jc) is at
// rlo * 32 + clo + 30 jr + q; the 8-bit sx / 16-bit sy*w fields are picked with one byte permute each.
template <int NCLASS, class F>
PG2_DEV uint32_t tile_texel_index(const F& f, const RowDesc& rd, const ColDesc& cd, uint32_t q) {
    const uint32_t jr = q >> 1;
    const uint32_t w = f.cell[(rd.rw >> 16) + (cd.cw & 31u) + jr * 30u + q];   // rw >> 16 = rlo * MAX_WIN
    const uint32_t cls = NCLASS > 1 ? (w >> 29) & 1u : 0u;
    const uint32_t sx = byte_perm(cd.csx, 0u, 0x4440u | (cls * 2u + (q & 1u)));
    const uint32_t syw = byte_perm(cls ? rd.syw[1] : rd.syw[0], 0u, 0x4410u + jr * 0x22u);
    return (w & CELL_OFFSET_MASK) + syw + sx;
}

// clear -> pre -> tiles of one pixel in reference (bottom-up) order with full blending: the path of pixels whose
// top-most layer is translucent, and of every pixel of a `wide` frame (more than two tiles cover a column / row,
// or a background the tables do not describe), which walks the covering ranges with the axes themselves.
template <class G, class F>
PG2_DEV_COLD uint32_t shade_base_ordered(const F& f, const uint32_t* __restrict__ atlas, int X, int Y) {
    uint32_t color = 0u, texel;   // SDL_RenderClear(0,0,0,255)
    for (int k = 0; k < f.npre; k++) {
        const FastBlit fb = f.fpre[k];
        BlitRot rot{ 0.0, 1.0, 0, 0 };
        if (!(fb.flags & 4u) && fast_texel<false>(fb, &rot, atlas, X, Y, &texel)) color = blend_packed(color, texel, fb.flags & 1u, fb.alpha_mod);
    }
    if (!f.wide) {
        const RowDesc rd = f.rowd[Y];
        const ColDesc cd = load_col(f, X);
        uint32_t p = tile_candidates<G::TILE_CLASSES>(rd, cd.cw);
        for (uint32_t q = 0; q < 4u; q++)
            if (p >> q & 1u) {
                texel = __ldg(atlas + tile_texel_index<G::TILE_CLASSES>(f, rd, cd, q));
                color = blend_packed(color, texel, 1u, 255u);
            }
        return color;
    }
    const int rlo = f.cov_lo[OBS_W + Y], rhi = f.cov_hi[OBS_W + Y], clo = f.cov_lo[X], chi = f.cov_hi[X];
    for (int ry = rlo; ry <= rhi; ry++)
        for (int cx = clo; cx <= chi; cx++) {
            const uint32_t w = f.cell[ry * MAX_WIN + cx];
            if (!w) continue;
            const uint32_t cls = (w >> 29) & 1u;
            const Axis& ax = f.col[cls][cx];
            const Axis& ay = f.row[cls][ry];
            if (!ax.visible || !ay.visible) continue;
            if ((unsigned)(X - ax.d0) >= (unsigned)ax.dlen || (unsigned)(Y - ay.d0) >= (unsigned)ay.dlen) continue;
            int sx = axis_sample(ax, X, false), sy = axis_sample(ay, Y, false);
            texel = __ldg(atlas + (w & CELL_OFFSET_MASK) + (uint32_t)sy * (uint32_t)f.class_w[cls] + (uint32_t)sx);
            color = blend_packed(color, texel, 1u, 255u);
        }
    return color;
}

// A pixel whose top-most tile texel turned out transparent: keep walking its candidates top-down (then the
// background); the first opaque texel decides, a translucent one hands the pixel to shade_base_ordered.
template <class G, class F>
PG2_DEV_COLD uint32_t shade_base_continue(const F& f, const uint32_t* __restrict__ atlas, int X, int Y, uint32_t p) {
    const RowDesc rd = f.rowd[Y];
    const ColDesc cd = load_col(f, X);
    while (p) {
        const uint32_t q = bfind(p);
        p &= ~(1u << q);
        const uint32_t texel = __ldg(atlas + tile_texel_index<G::TILE_CLASSES>(f, rd, cd, q));
        const uint32_t a = texel >> 24;
        if (a == 255u) return texel;
        if (a != 0u) return shade_base_ordered<G>(f, atlas, X, Y);
    }
    if ((rd.pre_row | cd.pre_sx) < 0) return 0u;
    const uint32_t texel = __ldg(atlas + (uint32_t)rd.pre_row + (uint32_t)cd.pre_sx);
    const uint32_t a = texel >> 24;
    if (a == 255u) return texel;
    return a ? shade_base_ordered<G>(f, atlas, X, Y) : 0u;
}

// Base pass of one band: clear + background + tile layer of 8 rows, packed into the warp's band buffer.
template <class G, class F>
PG2_DEV void raster_band_base(F& f, const uint32_t* __restrict__ atlas, int band, int lane, uint8_t* buf) {
    constexpr int NCLASS = G::TILE_CLASSES;
    const bool wide = f.wide != 0;
    for (int it = 0; it < BAND_ROWS / 2; it++)
        for (int l = lane; l < 32; l += WARP_LANES) {
            const int Y = band * BAND_ROWS + it * 2 + (l >> 4), X0 = (l & 15) * 4;
            uint32_t color[4];
            if (!wide) {
                const RowDesc rd = f.rowd[Y];
                uint32_t cand[4], texel[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    ColDesc cd;   // col_slot(X0 + i) = (l & 15) + 16 * i
                    cd.cw = f.col_cw[(l & 15) + 16 * i]; cd.csx = f.col_csx[(l & 15) + 16 * i]; cd.pre_sx = f.col_pre[(l & 15) + 16 * i];
                    const uint32_t p = G::HAS_TILES ? tile_candidates<NCLASS>(rd, cd.cw) : 0u;
                    cand[i] = p;
                    const uint32_t tidx = G::HAS_TILES ? tile_texel_index<NCLASS>(f, rd, cd, bfind(p | 1u)) : 0u;
                    const bool bg_ok = (rd.pre_row | cd.pre_sx) >= 0;
                    const uint32_t idx = p ? tidx : (uint32_t)rd.pre_row + (uint32_t)cd.pre_sx;
                    texel[i] = (p || bg_ok) ? __ldg(atlas + idx) : 0xff000000u;   // nothing there: the clear colour
                }
#pragma unroll
                for (int i = 0; i < 4; i++) color[i] = texel[i];
                // opaque-copy textures carry A = 255 in the atlas: ONE test covers the common case of four opaque texels
                if (((texel[0] & texel[1] & texel[2] & texel[3]) >> 24) != 255u) {
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const uint32_t a = texel[i] >> 24;
                        if (a != 255u) {   // transparent: next candidate below; translucent: blend in reference order
                            if (a != 0u) color[i] = shade_base_ordered<G>(f, atlas, X0 + i, Y);
                            else color[i] = cand[i] ? shade_base_continue<G>(f, atlas, X0 + i, Y, cand[i] & ~(1u << bfind(cand[i]))) : 0u;
                        }
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++) color[i] = shade_base_ordered<G>(f, atlas, X0 + i, Y);
            }
            uint32_t* out = (uint32_t*)(buf + 3 * ((it * 2 + (l >> 4)) * OBS_W + X0));
            out[0] = byte_perm(color[0], color[1], 0x4210u);
            out[1] = byte_perm(color[1], color[2], 0x5421u);
            out[2] = byte_perm(color[2], color[3], 0x6542u);
        }
}

// One post blit onto the rows [Y0, Y0 + 8) (the band buffer): lanes = an 8x4 patch of the destination rectangle.
template <class F>
PG2_DEV void draw_blit_band(F& f, const uint32_t* __restrict__ atlas, int k, int Y0, int lane, uint8_t* buf) {
    const FastBlit fb = f.fpost[k];
    const BlitRot* rot = &f.post_rot[F::ROTATES ? k : 0];
    int x0 = fb.x0, y0 = fb.y0, x1 = fb.x0 + fb.w - 1, y1 = fb.y0 + fb.h - 1;
    if (F::ROTATES && (fb.flags & 2u)) {   // bounding box of the rotated rect (rotated_extents)
        const int cx = fb.x0 + fb.w / 2, cy = fb.y0 + fb.h / 2;
        x0 = cx - rot->ex; x1 = cx + rot->ex; y0 = cy - rot->ey; y1 = cy + rot->ey;
    }
    x0 = max(x0, 0); x1 = min(x1, OBS_W - 1); y0 = max(y0, Y0); y1 = min(y1, Y0 + BAND_ROWS - 1);
    const uint32_t blend = fb.flags & 1u, alpha_mod = fb.alpha_mod;
    for (int yb = y0; yb <= y1; yb += 4)
        for (int xb = x0; xb <= x1; xb += 8)
            for (int l = lane; l < 32; l += WARP_LANES) {
                const int X = xb + (l & 7), Y = yb + (l >> 3);
                uint32_t texel;
                if (X <= x1 && Y <= y1 && fast_texel<F::ROTATES>(fb, rot, atlas, X, Y, &texel)) {
                    const uint32_t a = layer_alpha(texel, blend, alpha_mod);
                    uint8_t* px = buf + 3 * ((Y - Y0) * OBS_W + X);
                    if (a == 255u) { px[0] = (uint8_t)texel; px[1] = (uint8_t)(texel >> 8); px[2] = (uint8_t)(texel >> 16); }
                    else if (a != 0u) {
                        uint32_t r = px[0], g = px[1], b = px[2];
                        blend_texel(r, g, b, texel, blend, alpha_mod);
                        px[0] = (uint8_t)r; px[1] = (uint8_t)g; px[2] = (uint8_t)b;
                    }
                }
            }
    __syncwarp();
}

// The finished band leaves the SM as one 1 536-byte TMA bulk store shared -> global (async proxy), issued by the warp
// that drew it; the warp waits for the copy to have READ its band buffer right before it draws into the buffer again,
// so the drain overlaps the ticket / descriptor work in between.
PG2_DEV void band_store(uint8_t* __restrict__ dst, const uint8_t* buf, int lane) {
#ifdef PG2_HOSTSIM
    memcpy(dst, buf, BAND_BYTES);
#else
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy
    __syncwarp();
    if (lane == 0) {
        uint32_t src = (uint32_t)__cvta_generic_to_shared(buf);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "n"(BAND_BYTES) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
#endif
}

// Draw the frame: warps take bands from a ticket counter (a warp that drew cheap bands simply takes more of them),
// per band: base pass, post blits in submission order, bulk store.
// cache_img != nullptr (G::STATIC_VIEW): the env's base image; f.reuse says whether to load it or to (draw and) store it.
template <class G, class F>
PG2_DEV_NOINLINE void frame_rasterise(F& f, const uint32_t* __restrict__ atlas, uint8_t* __restrict__ dst, uint8_t* __restrict__ cache_img = nullptr) {
    const int lane = threadIdx.x % WARP_LANES;
    uint8_t* buf = f.band_rgb[threadIdx.x / WARP_LANES % (RENDER_THREADS / 32)];
    const int npost = f.npost;
    for (;;) {
        int band = 0;
        if (lane == 0) { band = smem_atomic_inc(&f.next_band); frame_store_wait(); }   // the previous band has left the buffer
        band = warp_bcast(band);
        if (band >= NUM_BANDS) break;
        if (cache_img != nullptr && f.reuse) {
            __syncwarp();
            band_from_cache(buf, cache_img + band * BAND_BYTES, lane);
        } else {
            raster_band_base<G>(f, atlas, band, lane, buf);
            if (cache_img != nullptr) { __syncwarp(); band_to_cache(buf, cache_img + band * BAND_BYTES, lane); }
        }
        __syncwarp();
        for (int base = 0; base < npost; base += 32) {
            uint32_t m = 0u;
            for (int l = lane; l < 32; l += WARP_LANES) {
                const int k = base + l;
                m |= lane_ballot(k < npost && (f.bandmask[k < npost ? k : 0] >> band & 1u), l);
            }
            while (m) {
                const int k = base + __ffs(m) - 1;
                m &= m - 1u;
                draw_blit_band(f, atlas, k, band * BAND_ROWS, lane, buf);
            }
        }
        band_store(dst + band * BAND_BYTES, buf, lane);
    }
}

// Per-CTA table of the game's tile textures (ids < MAX_TILE_TEX) as window cell words, filled once before the first frame.
template <class G, class F>
PG2_DEV void frame_init_tiletex(F& f, const TexInfo* __restrict__ tex) {
    for (int t = threadIdx.x; t < MAX_TILE_TEX; t += blockDim.x) {
        uint32_t w = 0u;
        if (t < G::NUM_TEX) {
            const TexInfo ti = tex[t];
            const uint32_t cls = (G::TILE_CLASSES > 1) ? (uint32_t)G::tile_class((uint32_t)t) : 0u;
            w = (ti.offset & CELL_OFFSET_MASK) | (ti.blend ? 1u << 28 : 0u) | cls << 29 | CELL_PRESENT;
        }
        f.tileword[t] = w;
    }
    for (int k = threadIdx.x; k < 2; k += blockDim.x) f.rowmask[k][MAX_WIN] = 0u;   // the always-empty tile row
    __syncthreads();
}

}  // namespace pg2
