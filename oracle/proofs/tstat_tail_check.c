/* tstat_tail_check.c -- TEST INFRASTRUCTURE. Validates the guarded shortcut for the last step of the
 * t-statistic (events.c:360)   t = (float)( fabs((double)delta) / sqrt((double)scaled) )
 * used by the CUDA fast path: y0 ~ 1/sqrt(scaled) to ~22 bits (MUFU.RSQ on the GPU; emulated here by a float
 * reciprocal square root perturbed by up to +-4 ulp), one third-order correction in double, q = |delta|*y,
 * accept (float)q unless q lies within 2^-44 (relative) of a float rounding midpoint or outside the normal float
 * range; otherwise the caller falls back to the IEEE sqrt + division. Reports mismatches among accepted values.
 * Build: gcc -O2 -mfma -ffp-contract=off -fopenmp tstat_tail_check.c -o tstat_tail_check -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static inline float f_from(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }
static inline uint32_t f_bits(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }
static inline uint64_t d_bits(double f) { uint64_t b; memcpy(&b, &f, 8); return b; }
static inline uint64_t splitmix(uint64_t *s) {
    uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

/* returns 1 and *out when the shortcut accepts */
static inline int tail_fast(float delta, float scaled, float y0f, float *out) {
    const double c = (double)scaled, y0 = (double)y0f;
    const double t = c * y0;
    const double e = fma(-t, y0, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double ye = y0 * e;
    const double y = fma(ye, p, y0);
    const double q = fabs((double)delta) * y;
    const uint64_t b = d_bits(q);
    const uint32_t lo29 = (uint32_t)(b & 0x1fffffffull);
    const uint32_t ex = (uint32_t)(b >> 52) & 0x7ff;
    /* the guard of tail() in sigtk_b200/csrc/walk_core.cuh: 2^-126 <= q < 2^126, the 29 bits below float
     * precision not within 512 of the rounding midpoint, scaled not tiny (the GPU seed flushes denormals) */
    const int in_range = ex >= 1023 - 126 && ex < 1023 + 126;
    const int off_mid = (uint32_t)(lo29 - (0x10000000u - 512u)) >= 1024u;
    if (!(in_range && off_mid && scaled >= 1.0e-30f)) return 0;
    *out = (float)q;
    return 1;
}

int main(void) {
    uint64_t bad = 0, rejected = 0, total = 0;
#pragma omp parallel for reduction(+ : bad, rejected, total)
    for (int t = 0; t < 64; t++) {
        uint64_t s = 777ull * (t + 1);
        for (uint64_t i = 0; i < (1ull << 24); i++) {
            const uint64_t r1 = splitmix(&s), r2 = splitmix(&s), r3 = splitmix(&s);
            /* scaled: positive float, exponent 2^-149 .. 2^60 ; delta: any sign, exponent 2^-60 .. 2^60 or zero */
            uint32_t cb = (uint32_t)(r1 & 0x7fffff) | ((uint32_t)(1 + (r1 >> 23) % 187) << 23);
            if ((r1 >> 40) % 1000 == 0) cb = (uint32_t)(r1 >> 41) & 0x7fffff;          /* denormal scaled */
            if (cb == 0) cb = 1;
            uint32_t db = (uint32_t)(r2 & 0x807fffff) | ((uint32_t)(67 + (r2 >> 32) % 121) << 23);
            if ((r2 >> 50) % 1000 == 0) db = 0;
            const float scaled = f_from(cb), delta = f_from(db);
            float y0 = (float)(1.0 / sqrt((double)scaled));
            int pert = (int)(r3 % 9) - 4;
            y0 = f_from(f_bits(y0) + pert);
            const float ref = (float)(fabs((double)delta) / sqrt((double)scaled));
            float got;
            total++;
            if (!tail_fast(delta, scaled, y0, &got)) { rejected++; continue; }
            if (f_bits(got) != f_bits(ref)) bad++;
        }
    }
    printf("tstat tail: %llu values, %llu rejected by the guard (%.3g), %llu mismatches among accepted\n",
           (unsigned long long)total, (unsigned long long)rejected, (double)rejected / (double)total,
           (unsigned long long)bad);
    return bad != 0;
}
