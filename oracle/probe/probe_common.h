/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 * Read-only probes into the reference's process globals, linked into oracle/_ref/lib<Game>.so
 * next to the unmodified reference sources so that parity tests can compare more than pixels:
 * the MT19937 position/state (RNG stream parity), the complete tile map (level layout parity)
 * and a few floats of game state. Nothing here writes to the reference's state — except pg2o_set_easy_mode (coinrun,
 * climber), which flips the generator's compile-time Config::easy_mode before cenv_make for the distribution-mode tests. */
#pragma once
#include <random>
#include <sstream>
#include <stdint.h>

extern std::mt19937 rng;   /* games/<g>/<g>.cpp:34 */

extern "C" {
/* 624 state words + position, through the engine's textual serialisation (operator<<):
 * libstdc++ writes the state rotated so that the next word to be tempered comes first?  No —
 * it writes _M_x[0..623] then _M_p. */
void pg2o_rng_state(uint32_t* out625) {
    std::ostringstream os;
    os << rng;
    std::istringstream is(os.str());
    for (int i = 0; i < 625; i++) { unsigned long v; is >> v; out625[i] = (uint32_t)v; }
}
}
