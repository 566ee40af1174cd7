#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
static inline float div5(float x, float h, float y) { // y = RN(1/h)
    float q0 = x * y;
    float r0 = fmaf(-q0, h, x);
    float q1 = fmaf(r0, y, q0);
    float r1 = fmaf(-q1, h, x);
    return fmaf(r1, y, q1);
}
static inline float div3(float x, float h, float y) {
    float q0 = x * y;
    float r0 = fmaf(-q0, h, x);
    return fmaf(r0, y, q0);
}
static uint64_t s = 88172645463325252ull;
static inline uint64_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
int main() {
    float hs[] = {0.2f, 0.8f, 0.05f, 0.1f, 0.3333333f, 1.0f, 0.0123f, 0.7f, 1.9999999f, 1.0000001f};
    for (int k = 0; k < 10; ++k) {
        float h = hs[k], y = 1.0f / h;
        long bad5 = 0, bad3 = 0, badr3 = 0, n = 400000000L;
        for (long i = 0; i < n; ++i) {
            uint32_t b = (uint32_t)rnd();
            float x; memcpy(&x, &b, 4);
            if (!isfinite(x) || fabsf(x) > 1e30f || (fabsf(x) < 1e-30f)) continue;
            float t = x / h;
            float a = div5(x, h, y), c = div3(x, h, y);
            if (a != t) bad5++;
            if (c != t) { bad3++; if (roundf(c) != roundf(t)) badr3++; }
        }
        printf("h=%.9g: div5 mismatches %ld, div3 mismatches %ld (round differs %ld)\n", h, bad5, bad3, badr3);
    }
}
