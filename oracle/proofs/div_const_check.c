/* div_const_check.c -- TEST INFRASTRUCTURE. Validates the arithmetic shortcuts used by the CUDA fast path
 * (sigtk_b200/csrc/walk_core.cuh) against IEEE division on the CPU:
 *   float : q = a / w  vs  q0 = a*r; e = fmaf(-w, q0, a); q = fmaf(e, r, q0)   EXHAUSTIVELY over all 2^32 floats
 *   double: same sequence on 4e9 random doubles per divisor in the magnitude range the path produces
 * for the window lengths w in {3, 6, 7, 14} (events.c:43-54).
 * Build: gcc -O2 -mfma -ffp-contract=off -fopenmp div_const_check.c -o div_const_check -lm
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

int main(void) {
    const int ws[4] = {3, 6, 7, 14};
    int rc = 0;
    for (int k = 0; k < 4; k++) {
        const float wf = (float)ws[k];
        const float rf = 1.0f / wf;
        uint64_t bad = 0;
        uint32_t lo_bad_exp = 0, hi_bad_exp = 0, min_bad_abs = 0xffffffffu, max_bad_abs = 0;
#pragma omp parallel for reduction(+ : bad) reduction(min : min_bad_abs) reduction(max : max_bad_abs)
        for (uint64_t b = 0; b < (1ull << 32); b++) {
            const float a = f_from((uint32_t)b);
            if (isnan(a) || isinf(a)) continue;
            const float ref = a / wf;
            const float q0 = a * rf;
            const float e = fmaf(-wf, q0, a);
            const float q = fmaf(e, rf, q0);
            if (f_bits(q) != f_bits(ref)) {
                bad++;
                const uint32_t abs = (uint32_t)b & 0x7fffffffu;
                if (abs < min_bad_abs) min_bad_abs = abs;
                if (abs > max_bad_abs) max_bad_abs = abs;
            }
        }
        (void)lo_bad_exp; (void)hi_bad_exp;
        printf("float  w=%2d: %llu mismatches of 2^32", ws[k], (unsigned long long)bad);
        if (bad) printf("  (|a| bit patterns from 0x%08x to 0x%08x, i.e. %g .. %g)", min_bad_abs, max_bad_abs,
                        f_from(min_bad_abs), f_from(max_bad_abs));
        printf("\n");
    }
    for (int k = 0; k < 4; k++) {
        const double wd = (double)ws[k];
        const double rd = 1.0 / wd;
        uint64_t bad = 0;
#pragma omp parallel for reduction(+ : bad)
        for (int t = 0; t < 64; t++) {
            uint64_t s = 0x1234567ull * (t + 1) + ws[k];
            for (uint64_t i = 0; i < (1ull << 26); i++) {
                const uint64_t m = splitmix(&s);
                /* random sign, random 52-bit mantissa, exponent in [-200, 200] */
                const uint64_t ex = 1023 - 200 + (splitmix(&s) % 401);
                uint64_t bits = (m & 0x800fffffffffffffull) | (ex << 52);
                double a; memcpy(&a, &bits, 8);
                const double ref = a / wd;
                const double q0 = a * rd;
                const double e = fma(-wd, q0, a);
                const double q = fma(e, rd, q0);
                if (d_bits(q) != d_bits(ref)) bad++;
            }
        }
        printf("double w=%2d: %llu mismatches of %llu random\n", ws[k], (unsigned long long)bad,
               (unsigned long long)(64ull << 26));
        if (bad) rc = 1;
    }
    return rc;
}
