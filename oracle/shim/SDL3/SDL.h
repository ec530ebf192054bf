/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * Minimal stand-in for the 17 SDL3 / SDL3_image symbols the reference games use
 * (SURVEY.md §2.3), so that the UNMODIFIED reference sources under
 * /root/reference/games/<g>/ compile into oracle/_ref/lib<Game>.so.
 * SDL3 itself is an un-vendored, un-pinned third-party dependency of the
 * reference (games/coinrun/CMakeLists.txt:20-21 `find_package(SDL3 REQUIRED)`);
 * its one draw primitive, SDL_RenderTextureRotated, is restated by the canonical
 * CPU rasteriser in oracle/raster.c (spec: DESIGN.md "Rasteriser specification").
 *
 * <stdlib.h>/<math.h> are included BY NAME on purpose: with
 * libstdc++ that is what brings the float overloads of abs() into the global
 * namespace, which chaser/jumper logic relies on.
 */
#pragma once
#include <stdlib.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <assert.h>
#include <time.h>

#ifdef __cplusplus
#include <type_traits>
static_assert(std::is_same<decltype(abs(0.5f)), float>::value, "abs(float) must resolve to the float overload (SURVEY Q16)");
extern "C" {
#endif

#define SDL_LIL_ENDIAN 1234
#define SDL_BIG_ENDIAN 4321
#define SDL_BYTEORDER SDL_LIL_ENDIAN
#define SDL_INIT_VIDEO 0x20u
#define SDL_LOG_CATEGORY_APPLICATION 0
#define SDL_LOG_PRIORITY_INFO 3

typedef uint8_t Uint8;
typedef uint32_t Uint32;

typedef struct SDL_FRect { float x, y, w, h; } SDL_FRect;
typedef struct SDL_Rect { int x, y, w, h; } SDL_Rect;
typedef struct SDL_FPoint { float x, y; } SDL_FPoint;

typedef enum { SDL_FLIP_NONE = 0, SDL_FLIP_HORIZONTAL = 1, SDL_FLIP_VERTICAL = 2 } SDL_RendererFlip;

typedef struct SDL_Surface {
    int w, h, pitch;
    void* pixels;      /* RGBA8, byte order R,G,B,A */
    int has_alpha;     /* source image carried an alpha channel / tRNS */
    int owns_pixels;
} SDL_Surface;

typedef struct SDL_Renderer {
    SDL_Surface* target;
    Uint8 r, g, b, a;
} SDL_Renderer;

typedef struct SDL_Texture {
    int w, h;
    int blend;         /* 1 = SDL_BLENDMODE_BLEND (alpha textures), 0 = copy */
    Uint8 alpha_mod;
    const uint8_t* rgba;
} SDL_Texture;

int SDL_Init(Uint32 flags);
void SDL_LogSetPriority(int category, int priority);
Uint32 SDL_GetPixelFormatEnumForMasks(int bpp, Uint32 r, Uint32 g, Uint32 b, Uint32 a);
SDL_Surface* SDL_CreateSurface(int w, int h, Uint32 format);
void SDL_DestroySurface(SDL_Surface* s);
SDL_Renderer* SDL_CreateSoftwareRenderer(SDL_Surface* s);
void SDL_DestroyRenderer(SDL_Renderer* r);
SDL_Texture* SDL_CreateTextureFromSurface(SDL_Renderer* r, SDL_Surface* s);
void SDL_DestroyTexture(SDL_Texture* t);
int SDL_SetTextureAlphaMod(SDL_Texture* t, Uint8 alpha);
int SDL_SetRenderDrawColor(SDL_Renderer* r, Uint8 cr, Uint8 cg, Uint8 cb, Uint8 ca);
int SDL_RenderClear(SDL_Renderer* r);
int SDL_RenderTextureRotated(SDL_Renderer* r, SDL_Texture* t, const SDL_FRect* src, const SDL_FRect* dst,
                             const double angle, const SDL_FPoint* center, const SDL_RendererFlip flip);
int SDL_LockSurface(SDL_Surface* s);
void SDL_UnlockSurface(SDL_Surface* s);

#ifdef __cplusplus
}
#endif
