/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 * Process-wide cache of decoded textures from the packed asset blob, built once as
 * oracle/_ref/libpg2o_assets.so so that many loaded copies of a reference game library
 * (one copy per environment: the reference keeps all state in globals) share one decode. */
#include "assets_blob.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <mutex>
#include <string>
#include <zlib.h>

extern "C" {

struct Cached { uint8_t* px; int w, h, has_alpha; };
static std::map<std::string, Cached> g_cache;
static std::mutex g_mu;

static uint8_t* blob_load_uncached(const char* blob_path, const char* name, int* w, int* h, int* has_alpha);

uint8_t* pg2o_blob_load(const char* blob_path, const char* name, int* w, int* h, int* has_alpha) {
    std::lock_guard<std::mutex> lk(g_mu);
    std::string key = std::string(blob_path) + "|" + name;
    auto it = g_cache.find(key);
    if (it == g_cache.end()) {
        Cached c{};
        c.px = blob_load_uncached(blob_path, name, &c.w, &c.h, &c.has_alpha);
        if (!c.px) return nullptr;
        it = g_cache.emplace(key, c).first;
    }
    *w = it->second.w; *h = it->second.h; *has_alpha = it->second.has_alpha;
    return it->second.px;
}

/* ---- blob reader -------------------------------------------------------------------- */
struct BlobEntry { char name[120]; uint32_t w, h, channels, reserved; uint64_t offset, zsize; };

static uint8_t* blob_load_uncached(const char* blob_path, const char* name, int* w, int* h, int* has_alpha) {
    FILE* f = fopen(blob_path, "rb");
    if (!f) return nullptr;
    char magic[8]; uint32_t ver = 0, count = 0;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "PG2ASSET", 8) != 0 || fread(&ver, 4, 1, f) != 1 || fread(&count, 4, 1, f) != 1) { fclose(f); return nullptr; }
    uint8_t* out = nullptr;
    for (uint32_t i = 0; i < count; i++) {
        BlobEntry e;
        if (fread(&e, sizeof(e), 1, f) != 1) break;
        if (strncmp(e.name, name, sizeof(e.name)) != 0) continue;
        uint8_t* z = (uint8_t*)malloc(e.zsize);
        fseek(f, (long)e.offset, SEEK_SET);
        if (fread(z, 1, e.zsize, f) != e.zsize) { free(z); break; }
        uLongf raw_len = (uLongf)e.w * e.h * e.channels;
        uint8_t* raw = (uint8_t*)malloc(raw_len);
        if (uncompress(raw, &raw_len, z, (uLong)e.zsize) != Z_OK) { free(z); free(raw); break; }
        free(z);
        size_t n = (size_t)e.w * e.h;
        out = (uint8_t*)malloc(n * 4);
        if (e.channels == 4) memcpy(out, raw, n * 4);
        else for (size_t p = 0; p < n; p++) { out[4 * p] = raw[3 * p]; out[4 * p + 1] = raw[3 * p + 1]; out[4 * p + 2] = raw[3 * p + 2]; out[4 * p + 3] = 255; }
        free(raw);
        *w = (int)e.w; *h = (int)e.h; *has_alpha = (e.channels == 4);
        break;
    }
    fclose(f);
    return out;
}

} /* extern "C" */
