// Is  q' = fma(fma(-q0, b, a), y, q0)  with  y = RN(1/b),  q0 = RN(a*y)  equal to RN(a/b) for float32?
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline float fast_div(float a, float b, float y) {
    float q0 = a * y;
    float r = fmaf(-q0, b, a);
    return fmaf(r, y, q0);
}
static uint64_t s = 88172645463325252ull;
static inline uint64_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static inline float bits2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
int main() {
    long bad = 0, n = 0, bad_allones = 0;
    // 1. b random in [2^-20, 2^20], a random in [2^-30, 2^30] (signs too)
    for (long i = 0; i < 400000000L; ++i) {
        uint32_t mb = rnd() & 0x7FFFFF, ma = rnd() & 0x7FFFFF;
        int eb = 127 - 20 + (int)(rnd() % 41), ea = 127 - 30 + (int)(rnd() % 61);
        float b = bits2f(((uint32_t)eb << 23) | mb), a = bits2f(((uint32_t)ea << 23) | ma);
        if (rnd() & 1) a = -a;
        float y = 1.0f / b;
        float q = a / b, f = fast_div(a, b, y);
        ++n;
        if (q != f) { ++bad; if (mb == 0x7FFFFF) ++bad_allones; if (bad < 10) printf("mismatch a=%a b=%a q=%a fast=%a mb=%06x\n", a, b, q, f, mb); }
    }
    printf("random: %ld / %ld mismatches (all-ones mantissa: %ld)\n", bad, n, bad_allones);
    // 2. b with all-ones mantissa
    bad = 0; n = 0;
    for (long i = 0; i < 50000000L; ++i) {
        float b = bits2f((127u << 23) | 0x7FFFFF);
        uint32_t ma = rnd() & 0x7FFFFF; int ea = 127 - 10 + (int)(rnd() % 21);
        float a = bits2f(((uint32_t)ea << 23) | ma);
        float y = 1.0f / b; ++n;
        if (a / b != fast_div(a, b, y)) ++bad;
    }
    printf("all-ones mantissa divisor: %ld / %ld mismatches\n", bad, n);
    // 3. integer-valued a (uint16 - bkg style): a = s - bkg for s in 0..65535, several (bkg, nrm)
    bad = 0; n = 0;
    for (int t = 0; t < 3000; ++t) {
        float bkg = (float)(rnd() % 4000) / 7.0f, nrm = 1.0f + (float)(rnd() % 300000) / 37.0f;
        float y = 1.0f / nrm;
        for (int sv = 0; sv < 65536; ++sv) {
            float a = (float)sv - bkg; ++n;
            if (a / nrm != fast_div(a, nrm, y)) ++bad;
        }
    }
    printf("uint16 inputs: %ld / %ld mismatches\n", bad, n);
    // 4. x / n with x in [0,1], n in (0,4]
    bad = 0; n = 0;
    for (long i = 0; i < 200000000L; ++i) {
        float x = (float)(rnd() & 0xFFFFFF) / 16777216.0f;
        float nn = 1e-3f + 4.0f * (float)(rnd() & 0xFFFFFF) / 16777216.0f;
        float y = 1.0f / nn; ++n;
        if (x / nn != fast_div(x, nn, y)) ++bad;
    }
    printf("unit-vector style: %ld / %ld mismatches\n", bad, n);
    return 0;
}
