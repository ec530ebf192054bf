// Plain types shared by host-only and device translation units.
#pragma once
#include <stdint.h>

namespace pg2 {

// Texture atlas: every texture a game can draw, decoded to RGBA8 (one u32 per texel,
// little-endian R,G,B,A) in ONE device allocation; `TexInfo` is the rect table. atlas[0] is one opaque black texel
// (the clear colour), textures follow.
struct TexInfo {
    uint32_t offset;   // first texel (u32 index into the atlas)
    uint16_t w, h;
    uint32_t blend;    // 1: alpha texture (SRC-over), 0: opaque copy
};

}  // namespace pg2
