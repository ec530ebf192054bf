/* TEST INFRASTRUCTURE — NOT PRODUCT CODE. See raster.h for scope and the
 * "parity unpinned" statement. Compile with -ffp-contract=off (no FMA). */
#include "raster.h"
#include <math.h>

/* ---- deterministic sin/cos in degrees -------------------------------------------- */
void pg2o_sincos_deg(double deg, double* s, double* c) {
    double r = fmod(deg, 360.0);            /* exact */
    if (r < 0.0) r = r + 360.0;
    int q = (int)((r + 45.0) / 90.0);       /* 0..4 */
    double t = r - (double)q * 90.0;        /* [-45, 45] */
    double x = t * 0.017453292519943295;    /* pi/180 */
    double x2 = x * x;
    /* Taylor series, Horner form, every operation individually rounded */
    double ps = -1.0 / 355687428096000.0;                 /* -1/17! */
    ps = ps * x2 + 1.0 / 1307674368000.0;                 /*  1/15! */
    ps = ps * x2 - 1.0 / 6227020800.0;                    /* -1/13! */
    ps = ps * x2 + 1.0 / 39916800.0;                      /*  1/11! */
    ps = ps * x2 - 1.0 / 362880.0;                        /* -1/9!  */
    ps = ps * x2 + 1.0 / 5040.0;                          /*  1/7!  */
    ps = ps * x2 - 1.0 / 120.0;                           /* -1/5!  */
    ps = ps * x2 - 1.0 / 6.0;                             /* -1/3!  */
    double s0 = x + x * (x2 * ps);
    double pc = 1.0 / 20922789888000.0;                   /*  1/16! */
    pc = pc * x2 - 1.0 / 87178291200.0;                   /* -1/14! */
    pc = pc * x2 + 1.0 / 479001600.0;                     /*  1/12! */
    pc = pc * x2 - 1.0 / 3628800.0;                       /* -1/10! */
    pc = pc * x2 + 1.0 / 40320.0;                         /*  1/8!  */
    pc = pc * x2 - 1.0 / 720.0;                           /* -1/6!  */
    pc = pc * x2 + 1.0 / 24.0;                            /*  1/4!  */
    pc = pc * x2 - 0.5;                                   /* -1/2!  */
    double c0 = 1.0 + x2 * pc;
    switch (q & 3) {
    case 0: *s = s0;  *c = c0;  break;
    case 1: *s = c0;  *c = -s0; break;
    case 2: *s = -s0; *c = -c0; break;
    default: *s = -c0; *c = s0; break;
    }
}

/* ---- blend ------------------------------------------------------------------------ */
static inline void blend_px(uint8_t* d, const uint8_t* t, int blend, uint8_t alpha_mod) {
    if (!blend) {           /* opaque texture: straight copy, alpha-mod ignored */
        d[0] = t[0]; d[1] = t[1]; d[2] = t[2]; d[3] = 255;
        return;
    }
    unsigned a = t[3];
    if (alpha_mod != 255) a = (a * alpha_mod) / 255;
    unsigned r = t[0], g = t[1], b = t[2];
    if (a < 255) { r = (r * a) / 255; g = (g * a) / 255; b = (b * a) / 255; }
    unsigned ia = 255 - a;
    d[0] = (uint8_t)(r + (ia * d[0]) / 255);
    d[1] = (uint8_t)(g + (ia * d[1]) / 255);
    d[2] = (uint8_t)(b + (ia * d[2]) / 255);
    d[3] = (uint8_t)(a + (ia * d[3]) / 255);
}

void pg2o_blit(uint8_t* target, int tw, int th, const pg2o_texture* tex,
               const float* src_xywh, const float* dst_xywh,
               double angle_deg, int flip, uint8_t alpha_mod) {
    /* 1. source rect: intersect (in float) with the texture bounds; the destination is
     *    NOT re-adjusted (SDL_RenderTexture semantics). */
    float sxf = 0.0f, syf = 0.0f, swf = (float)tex->w, shf = (float)tex->h;
    if (src_xywh) {
        float amin = src_xywh[0], amax = amin + src_xywh[2];
        if (amin < 0.0f) amin = 0.0f;
        if (amax > (float)tex->w) amax = (float)tex->w;
        sxf = amin; swf = amax - amin;
        amin = src_xywh[1]; amax = amin + src_xywh[3];
        if (amin < 0.0f) amin = 0.0f;
        if (amax > (float)tex->h) amax = (float)tex->h;
        syf = amin; shf = amax - amin;
    }
    int sx = (int)sxf, sy = (int)syf, sw = (int)swf, sh = (int)shf;
    if (sw <= 0 || sh <= 0) return;
    /* 2. destination rect truncated toward zero */
    int dx = (int)dst_xywh[0], dy = (int)dst_xywh[1], dw = (int)dst_xywh[2], dh = (int)dst_xywh[3];
    if (dw <= 0 || dh <= 0) return;
    /* 3. 16.16 fixed-point nearest sampling, starting at the pixel centre */
    uint32_t incx = (uint32_t)(((uint64_t)sw << 16) / (uint64_t)dw);
    uint32_t incy = (uint32_t)(((uint64_t)sh << 16) / (uint64_t)dh);

    if (angle_deg == 0.0) {
        for (int j = 0; j < dh; j++) {
            int ty = dy + j;
            if (ty < 0 || ty >= th) continue;
            int jj = (flip == 2) ? dh - 1 - j : j;
            int srcy = sy + (int)((incy / 2 + (uint32_t)jj * incy) >> 16);
            for (int i = 0; i < dw; i++) {
                int tx = dx + i;
                if (tx < 0 || tx >= tw) continue;
                int ii = (flip == 1) ? dw - 1 - i : i;
                int srcx = sx + (int)((incx / 2 + (uint32_t)ii * incx) >> 16);
                blend_px(target + 4 * (ty * tw + tx), tex->rgba + 4 * (srcy * tex->w + srcx), tex->blend, alpha_mod);
            }
        }
        return;
    }
    /* 4. rotation about the centre of the integer destination rect: every target pixel
     *    centre is mapped back into the un-rotated rect (nearest). */
    double sn, cs;
    pg2o_sincos_deg(angle_deg, &sn, &cs);
    double hw = (double)dw * 0.5, hh = (double)dh * 0.5;
    double cx = (double)dx + hw, cy = (double)dy + hh;
    for (int ty = 0; ty < th; ty++)
        for (int tx = 0; tx < tw; tx++) {
            double px = ((double)tx + 0.5) - cx;
            double py = ((double)ty + 0.5) - cy;
            double u = px * cs + py * sn;
            double v = py * cs - px * sn;
            double fu = floor(u + hw), fv = floor(v + hh);
            if (fu < 0.0 || fv < 0.0 || fu >= (double)dw || fv >= (double)dh) continue;
            int i = (int)fu, j = (int)fv;
            int ii = (flip == 1) ? dw - 1 - i : i;
            int jj = (flip == 2) ? dh - 1 - j : j;
            int srcx = sx + (int)((incx / 2 + (uint32_t)ii * incx) >> 16);
            int srcy = sy + (int)((incy / 2 + (uint32_t)jj * incy) >> 16);
            blend_px(target + 4 * (ty * tw + tx), tex->rgba + 4 * (srcy * tex->w + srcx), tex->blend, alpha_mod);
        }
}
