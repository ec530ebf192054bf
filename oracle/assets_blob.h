/* Reader for the packed asset blob (procgen2_b200/data/assets.bin). Shared FORMAT
 * definition only — the product has its own reader in procgen2_b200/csrc/assets.cpp;
 * this copy is used by the oracle shim's IMG_Load. TEST INFRASTRUCTURE. */
#ifndef PG2_ASSETS_BLOB_H
#define PG2_ASSETS_BLOB_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* Returns malloc'ed RGBA8 pixels (caller keeps forever), or NULL when `name` is absent. */
uint8_t* pg2o_blob_load(const char* blob_path, const char* name, int* w, int* h, int* has_alpha);
#ifdef __cplusplus
}
#endif
#endif
