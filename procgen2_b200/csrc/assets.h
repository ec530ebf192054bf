// Reader of the packed asset blob (procgen2_b200/pack_assets.py documents the layout) -> host
// RGBA8 texels + rect table, uploaded once per GPU as the device texture atlas. Replaces
// Asset_Texture::load (games/coinrun/common_assets.cpp:3-17: IMG_Load + SDL_CreateTextureFromSurface).
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include "pg2_types.h"

namespace pg2 {
bool load_textures(const char* blob_path, const char* const* names, int count, std::vector<TexInfo>* infos,
                   std::vector<uint32_t>* texels, std::string* err);
std::string default_assets_path();
}
