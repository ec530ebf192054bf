// TEST: pg2::glibc_sincosf (procgen2_b200/csrc/pg2_libm.cuh) == the host libm's sincosf, bit for
// bit, over (a) every float in a few dense windows, (b) a strided sweep of all finite floats,
// (c) the argument ranges the games produce. argv[1] = stride of the global sweep.
#define PG2_HOSTSIM 1
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../procgen2_b200/csrc/pg2_libm.cuh"

using namespace pg2;

static long bad = 0, total = 0;
static void check(uint32_t bits) {
    float y; memcpy(&y, &bits, 4);
    float s0, c0, s1, c1;
    sincosf(y, &s0, &c0);
    glibc_sincosf(y, &s1, &c1);
    total++;
    if (memcmp(&s0, &s1, 4) || memcmp(&c0, &c1, 4)) {
        if (isnan(s0) && isnan(s1) && isnan(c0) && isnan(c1)) return;
        if (bad < 10) printf("MISMATCH y=%a libm=(%a,%a) mine=(%a,%a)\n", y, s0, c0, s1, c1);
        bad++;
    }
}

int main(int argc, char** argv) {
    uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 1021u;
    for (uint64_t b = 0; b < 0x100000000ull; b += stride) check((uint32_t)b);
    // dense: [0.5, 16) positive and negative (bossfight / caveflyer angles), tiny values, the 120 boundary
    for (uint32_t b = 0x3f000000u; b < 0x41800000u; b += 7) { check(b); check(b | 0x80000000u); }
    for (uint32_t b = 0x39000000u; b < 0x39900000u; b += 3) check(b);
    for (uint32_t b = 0x42ef0000u; b < 0x42f10000u; b++) check(b);
    // atan2f: random bit patterns, game-range values, axis / sign / tiny / huge edge cases
    {
        uint64_t st = 0x9e3779b97f4a7c15ull;
        auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (uint32_t)(st >> 16); };
        auto chk2 = [&](float y, float x) {
            float a = atan2f(y, x), b = glibc_atan2f(y, x);
            total++;
            if (memcmp(&a, &b, 4) && !(isnan(a) && isnan(b))) { if (bad < 10) printf("MISMATCH atan2f(%a,%a) libm=%a mine=%a\n", y, x, a, b); bad++; }
        };
        long n2 = 40000000L / (stride > 1000 ? 1 : 1);
        for (long i = 0; i < n2; i++) {
            uint32_t a = rnd(), b = rnd();
            float y, x; memcpy(&y, &a, 4); memcpy(&x, &b, 4);
            chk2(y, x);
            chk2(((int)(a % 80001) - 40000) * 0.001f, ((int)(b % 80001) - 40000) * 0.001f);   // to_goal range
        }
        const float sp[] = { 0.0f, -0.0f, 1.0f, -1.0f, 1e-38f, -1e-38f, 1e38f, -1e38f, INFINITY, -INFINITY, 0.4375f, 0.6875f, 1.1875f, 2.4375f, 3.0e7f, 4.0e7f };
        for (float y : sp) for (float x : sp) chk2(y, x);
    }
    printf("%s %ld checked, %ld mismatches\n", bad ? "FAIL" : "OK", total, bad);
    return bad ? 1 : 0;
}
