/* TEST INFRASTRUCTURE — see SDL.h in this directory. */
#pragma once
#include "SDL.h"
#ifdef __cplusplus
extern "C" {
#endif
#define IMG_INIT_PNG 2
int IMG_Init(int flags);
/* Looks `path` ("assets/...png") up in the packed asset blob (procgen2_b200/data/assets.bin,
 * produced from the reference PNGs by procgen2_b200/pack_assets.py using PIL). */
SDL_Surface* IMG_Load(const char* path);
#ifdef __cplusplus
}
#endif
