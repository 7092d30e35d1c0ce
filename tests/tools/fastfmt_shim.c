/* fastfmt_shim.c -- TEST INFRASTRUCTURE: cli/fastfmt.h against glibc's snprintf (tests/test_fastfmt.py). */
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../cli/fastfmt.h"

static uint64_t splitmix(uint64_t *s) {
    uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
/* number of floats whose fmt_f6 text differs from "%f"; mode 0: random bit patterns, 1: pA-like values,
 * 2: exact ties and neighbours (k / 2^j around 6-decimal boundaries) */
uint64_t fastfmt_selftest(uint64_t seed, uint64_t n, int mode, float *first_bad) {
    uint64_t bad = 0, s = seed;
    char a[64], b[64];
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t r = splitmix(&s);
        float f;
        if (mode == 0) { uint32_t u = (uint32_t)r; memcpy(&f, &u, 4); }
        else if (mode == 1) f = (float)((double)(r >> 11) / 9007199254740992.0 * 400.0 - 50.0);
        else { f = (float)((double)((r >> 8) % 20000001u) / (double)(1u << (r % 25))); if (r & 1) f = -f; }
        *fmt_f6(a, f) = 0;
        snprintf(b, sizeof b, "%f", (double)f);
        if (strcmp(a, b)) { if (!bad && first_bad) *first_bad = f; bad++; }
    }
    return bad;
}
uint64_t fastfmt_int_selftest(uint64_t seed, uint64_t n) {
    uint64_t bad = 0, s = seed;
    char a[64], b[64];
    for (uint64_t i = 0; i < n; i++) {
        const int64_t v = (int64_t)splitmix(&s) >> (splitmix(&s) % 64);
        *fmt_i64(a, v) = 0;
        snprintf(b, sizeof b, "%ld", (long)v);
        if (strcmp(a, b)) bad++;
    }
    const int64_t edge[] = {0, -1, 1, INT64_MIN, INT64_MAX, -2147483648ll, 2147483647ll};
    for (unsigned k = 0; k < sizeof edge / sizeof edge[0]; k++) {
        *fmt_i64(a, edge[k]) = 0;
        snprintf(b, sizeof b, "%ld", (long)edge[k]);
        if (strcmp(a, b)) bad++;
    }
    return bad;
}
int fastfmt_one(float f, char *out) { char *e = fmt_f6(out, f); *e = 0; return (int)(e - out); }
