#include "assets.h"

#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <map>

namespace pg2 {

namespace {
struct Entry { char name[120]; uint32_t w, h, channels, reserved; uint64_t offset, zsize; };
static_assert(sizeof(Entry) == 152, "blob entry layout");
}

std::string default_assets_path() {
    const char* env = getenv("PG2_ASSETS");
    if (env && *env) return env;
    Dl_info info;
    if (dladdr((void*)&default_assets_path, &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        size_t slash = p.rfind('/');
        std::string dir = slash == std::string::npos ? "." : p.substr(0, slash);
        return dir + "/../data/assets.bin";
    }
    return "procgen2_b200/data/assets.bin";
}

bool load_textures(const char* blob_path, const char* const* names, int count, std::vector<TexInfo>* infos,
                   std::vector<uint32_t>* texels, std::string* err) {
    std::string path = (blob_path && *blob_path) ? std::string(blob_path) : default_assets_path();
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { *err = "cannot open asset blob '" + path + "'"; return false; }
    char magic[8]; uint32_t ver = 0, n = 0;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "PG2ASSET", 8) != 0 || fread(&ver, 4, 1, f) != 1 || fread(&n, 4, 1, f) != 1 || ver != 1) {
        fclose(f); *err = "bad asset blob header in '" + path + "'"; return false;
    }
    std::vector<Entry> entries(n);
    if (fread(entries.data(), sizeof(Entry), n, f) != n) { fclose(f); *err = "truncated asset blob"; return false; }
    std::map<std::string, int> index;
    for (uint32_t i = 0; i < n; i++) index[std::string(entries[i].name, strnlen(entries[i].name, sizeof(entries[i].name)))] = (int)i;
    infos->resize(count);
    texels->clear();
    texels->push_back(0xff000000u);   // atlas[0] = opaque black = the clear colour (what the rasteriser reads where nothing is drawn)
    std::vector<uint8_t> z, raw;
    for (int t = 0; t < count; t++) {
        auto it = index.find(names[t]);
        if (it == index.end()) { fclose(f); *err = std::string("texture '") + names[t] + "' missing from asset blob"; return false; }
        const Entry& e = entries[it->second];
        z.resize(e.zsize);
        raw.resize((size_t)e.w * e.h * e.channels);
        fseek(f, (long)e.offset, SEEK_SET);
        uLongf raw_len = (uLongf)raw.size();
        if (fread(z.data(), 1, e.zsize, f) != e.zsize || uncompress(raw.data(), &raw_len, z.data(), (uLong)e.zsize) != Z_OK || raw_len != raw.size()) {
            fclose(f); *err = std::string("cannot inflate texture '") + names[t] + "'"; return false;
        }
        TexInfo& ti = (*infos)[t];
        ti.offset = (uint32_t)texels->size(); ti.w = (uint16_t)e.w; ti.h = (uint16_t)e.h; ti.blend = (e.channels == 4);
        size_t px = (size_t)e.w * e.h;
        size_t base = texels->size();
        texels->resize(base + px);
        uint32_t* out = texels->data() + base;
        if (e.channels == 4) memcpy(out, raw.data(), px * 4);
        else for (size_t p = 0; p < px; p++) out[p] = (uint32_t)raw[3 * p] | (uint32_t)raw[3 * p + 1] << 8 | (uint32_t)raw[3 * p + 2] << 16 | 0xff000000u;
    }
    fclose(f);
    return true;
}

}  // namespace pg2
