/* TEST INFRASTRUCTURE — NOT PRODUCT CODE (only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may link or call this).
 *
 * Canonical CPU rasteriser: a plain-C restatement of the ONE draw primitive the
 * reference ever issues, SDL_RenderTextureRotated (call sites
 * games/<g>/renderer.cpp:78, renderer.cpp:97, games/jumper/jumper.cpp:489,500,508).
 * The arithmetic lives in SDL3's software renderer, an un-vendored and un-pinned
 * third-party dependency (no version anywhere in /root/reference), so PIXEL PARITY
 * IS "PARITY UNPINNED": this file is the written specification, restating the
 * published SDL software-blit algorithm (nearest-neighbour 16.16 fixed-point scaled
 * blit with centre-of-pixel start, integer SRC-over blend with /255 division) under
 * the assumptions listed in DESIGN.md "Rasteriser specification".
 */
#ifndef PG2_ORACLE_RASTER_H
#define PG2_ORACLE_RASTER_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int w, h;
    int blend;            /* 1: texture has an alpha channel -> SRC-over; 0: opaque copy */
    const uint8_t* rgba;  /* w*h*4, R,G,B,A */
} pg2o_texture;

/* Target is tw x th RGBA8 (R,G,B,A byte order), row-major.
 * src may be NULL (whole texture). angle in degrees (clockwise on screen, as SDL).
 * flip: 0 none, 1 horizontal, 2 vertical. */
void pg2o_blit(uint8_t* target, int tw, int th, const pg2o_texture* tex,
               const float* src_xywh, const float* dst_xywh,
               double angle_deg, int flip, uint8_t alpha_mod);

/* Deterministic sin/cos of an angle given in degrees (IEEE double mul/add only,
 * fixed operation order, no FMA) — mirrored operation by operation on the device. */
void pg2o_sincos_deg(double deg, double* s, double* c);

#ifdef __cplusplus
}
#endif
#endif
