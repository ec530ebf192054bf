/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 * Implementation of the SDL3 stand-in declared in SDL3/SDL.h: software "renderer"
 * state + dispatch of the single draw primitive to oracle/raster.c. */
#include "SDL3/SDL.h"
#include "SDL3/SDL_image.h"
#include "../raster.h"
#include "../assets_blob.h"

#include <stdio.h>
#include <string>


extern "C" {

int SDL_Init(Uint32) { return 0; }
void SDL_LogSetPriority(int, int) {}
Uint32 SDL_GetPixelFormatEnumForMasks(int, Uint32, Uint32, Uint32, Uint32) { return 1; }

SDL_Surface* SDL_CreateSurface(int w, int h, Uint32) {
    SDL_Surface* s = (SDL_Surface*)calloc(1, sizeof(SDL_Surface));
    s->w = w; s->h = h; s->pitch = 4 * w;
    s->pixels = calloc((size_t)w * h, 4);
    s->has_alpha = 1; s->owns_pixels = 1;
    return s;
}
void SDL_DestroySurface(SDL_Surface* s) {
    if (!s) return;
    if (s->owns_pixels) free(s->pixels);
    free(s);
}
SDL_Renderer* SDL_CreateSoftwareRenderer(SDL_Surface* s) {
    SDL_Renderer* r = (SDL_Renderer*)calloc(1, sizeof(SDL_Renderer));
    r->target = s; r->a = 255;
    return r;
}
void SDL_DestroyRenderer(SDL_Renderer* r) { free(r); }

SDL_Texture* SDL_CreateTextureFromSurface(SDL_Renderer*, SDL_Surface* s) {
    SDL_Texture* t = (SDL_Texture*)calloc(1, sizeof(SDL_Texture));
    t->w = s->w; t->h = s->h; t->blend = s->has_alpha; t->alpha_mod = 255;
    t->rgba = (const uint8_t*)s->pixels;   /* pixel storage is shared and never freed (s->owns_pixels == 0) */
    return t;
}
void SDL_DestroyTexture(SDL_Texture* t) { free(t); }
int SDL_SetTextureAlphaMod(SDL_Texture* t, Uint8 a) { t->alpha_mod = a; return 0; }
int SDL_SetRenderDrawColor(SDL_Renderer* r, Uint8 cr, Uint8 cg, Uint8 cb, Uint8 ca) { r->r = cr; r->g = cg; r->b = cb; r->a = ca; return 0; }
int SDL_RenderClear(SDL_Renderer* r) {
    uint8_t* p = (uint8_t*)r->target->pixels;
    size_t n = (size_t)r->target->w * r->target->h;
    for (size_t i = 0; i < n; i++) { p[4 * i] = r->r; p[4 * i + 1] = r->g; p[4 * i + 2] = r->b; p[4 * i + 3] = r->a; }
    return 0;
}

static long g_blit_count = 0;
long pg2o_blit_count() { return g_blit_count; }

int SDL_RenderTextureRotated(SDL_Renderer* r, SDL_Texture* t, const SDL_FRect* src, const SDL_FRect* dst,
                             const double angle, const SDL_FPoint*, const SDL_RendererFlip flip) {
    pg2o_texture tex{ t->w, t->h, t->blend, t->rgba };
    float d[4] = { dst->x, dst->y, dst->w, dst->h };
    float s[4];
    if (src) { s[0] = src->x; s[1] = src->y; s[2] = src->w; s[3] = src->h; }
    g_blit_count++;
    pg2o_blit((uint8_t*)r->target->pixels, r->target->w, r->target->h, &tex, src ? s : nullptr, d, angle, (int)flip, t->alpha_mod);
    return 0;
}
int SDL_LockSurface(SDL_Surface*) { return 0; }
void SDL_UnlockSurface(SDL_Surface*) {}

int IMG_Init(int flags) { return flags; }

SDL_Surface* IMG_Load(const char* path) {
    const char* discover = getenv("PG2O_DISCOVER");
    if (discover) {
        /* asset discovery mode (used once by pack_assets.py): log the path, return a
         * blank surface with the PNG's IHDR size, CWD must be the reference root. */
        FILE* f = fopen(path, "rb");
        if (!f) return nullptr;
        unsigned char hdr[24];
        size_t n = fread(hdr, 1, 24, f);
        fclose(f);
        if (n != 24) return nullptr;
        int w = (hdr[16] << 24) | (hdr[17] << 16) | (hdr[18] << 8) | hdr[19];
        int h = (hdr[20] << 24) | (hdr[21] << 16) | (hdr[22] << 8) | hdr[23];
        FILE* lg = fopen(discover, "a");
        if (lg) { fprintf(lg, "%s\n", path); fclose(lg); }
        SDL_Surface* s = SDL_CreateSurface(w, h, 0);
        s->owns_pixels = 0;
        return s;
    }
    const char* blob = getenv("PG2_ASSETS");
    if (!blob) blob = "procgen2_b200/data/assets.bin";
    int w, h, has_alpha;
    uint8_t* px = pg2o_blob_load(blob, path, &w, &h, &has_alpha);
    if (!px) { fprintf(stderr, "[oracle shim] asset '%s' not found in blob '%s'\n", path, blob); return nullptr; }
    SDL_Surface* s = (SDL_Surface*)calloc(1, sizeof(SDL_Surface));
    s->w = w; s->h = h; s->pitch = 4 * w; s->pixels = px; s->has_alpha = has_alpha; s->owns_pixels = 0;
    return s;
}

} /* extern "C" */
