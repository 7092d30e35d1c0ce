/* div_window_check.c -- TEST INFRASTRUCTURE. `sigtk prefix` divides the rolling-window sum (an exact integer in
 * [0, 1200 * 2000], jnn.c:20-50 on samples clamped to [0, 1200]) by the window length 2000 in float for every
 * position, three times. prefix.cu replaces the IEEE division by
 *     q0 = s * r;  e = fmaf(-2000, q0, s);  q = fmaf(e, r, q0),   r = 1.0f / 2000.0f
 * This program checks the two against each other for EVERY value the sum can take (and every integer up to 2^24,
 * the range in which the float sum is exact at all).
 * Build: gcc -O2 -mfma -ffp-contract=off div_window_check.c -o div_window_check -lm     exit status 1 on a mismatch */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

int main(void) {
    const float w = 2000.0f, r = 1.0f / 2000.0f;
    unsigned long bad = 0;
    for (uint32_t s = 0; s <= (1u << 24); s++) {
        const float a = (float)s;
        const float ref = a / w;
        const float q0 = a * r;
        const float e = fmaf(-w, q0, a);
        const float q = fmaf(e, r, q0);
        if (memcmp(&q, &ref, 4) != 0) {
            if (bad < 10) fprintf(stderr, "mismatch at s = %u: %.9g vs %.9g\n", s, q, ref);
            bad++;
        }
    }
    printf("{\"divisor\": 2000, \"values\": %u, \"mismatches\": %lu}\n", (1u << 24) + 1u, bad);
    return bad ? 1 : 0;
}
