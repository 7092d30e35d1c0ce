/* tstat_tail_check.c -- TEST INFRASTRUCTURE. Validates the guarded shortcut for the last step of the
 * t-statistic (events.c:360)   t = (float)( fabs((double)delta) / sqrt((double)scaled) )
 * used by the CUDA fast path: y0 ~ 1/sqrt(scaled) to ~22 bits (MUFU.RSQ on the GPU: PTX rsqrt.approx.f32, maximum
 * relative error 2^-22.4; emulated here by any float within that bound of the true value), ONE second-order correction in double,
 *     eh = 1/2 - (scaled/2) * y0^2,   q = |delta| * y0 * (1 + eh)        (4 double operations)
 * whose method error is 1.5 eps^2 <= 2^-44.2 relative (<= 445 ulp of a double). (float)q is accepted unless the 29
 * bits of q below float precision lie within 1024 of the rounding midpoint; otherwise the caller falls back to the
 * IEEE sqrt + division. No range check: the caller guarantees delta == 0 or 2^-84 <= |delta| <= 2^21 and
 * 1e-30 <= scaled <= 2^42 (walk_core.cuh, tail()), which is the domain sampled here. Reports mismatches among
 * accepted values.
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
    const double ch = (double)(scaled * 0.5f), y0 = (double)y0f;   /* scaled/2 is exact */
    const double th = ch * y0;
    const double eh = fma(-th, y0, 0.5);
    const double q0 = fabs((double)delta) * y0;
    const double q = fma(q0, eh, q0);
    const uint64_t b = d_bits(q);
    const uint32_t lo = (uint32_t)b;
    /* the guard of tail() in sigtk_b200/csrc/walk_core.cuh: the 29 bits below float precision not within 1024 of
     * the rounding midpoint (computed on the low word shifted up by 3, exactly as the kernel does) */
    const int off_mid = (uint32_t)(lo * 8u - ((0x10000000u - 1024u) << 3)) >= (2048u << 3);
    if (!off_mid) return 0;
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
            /* scaled: 2^-99 (< 1e-30) .. 2^43 ; delta: any sign, 2^-84 .. 2^21, or zero */
            uint32_t cb = (uint32_t)(r1 & 0x7fffff) | ((uint32_t)(28 + (r1 >> 23) % 143) << 23);
            if (f_from(cb) < 1.0e-30f) cb = f_bits(1.0e-30f);
            uint32_t db = (uint32_t)(r2 & 0x807fffff) | ((uint32_t)(43 + (r2 >> 32) % 106) << 23);
            if ((r2 >> 50) % 1000 == 0) db = 0;
            const float scaled = f_from(cb), delta = f_from(db);
            /* the seed: any float within 2^-22.4 (relative) of 1/sqrt(scaled) -- the bound PTX states for
             * rsqrt.approx.f32, and tools/microbench/rsqrt_err.cu measures on the GPU over all floats */
            const double u = ((double)(r3 >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0) * 1.80e-7; /* 2^-22.4 = 1.81e-7 */
            float y0 = (float)((1.0 / sqrt((double)scaled)) * (1.0 + u));
            if (fabs((double)y0 * sqrt((double)scaled) - 1.0) > 1.81e-7) y0 = (float)(1.0 / sqrt((double)scaled));
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
