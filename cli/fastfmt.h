/* fastfmt.h -- exact, allocation-free replacements for the printf conversions of the reference's output path
 * (src/cfunc.c:19-58, 90-99, 141): "%f" of a float (promoted to double), "%d", "%ld", "%s".
 *
 * glibc prints "%f" as the EXACT binary value rounded to 6 decimals, ties to even (default rounding mode). A
 * float is M * 2^E with M < 2^24, so value * 10^6 = M * 10^6 * 2^E is an integer (E >= 0) or a 44-bit integer
 * shifted right (E < 0): the rounding is done exactly in 64-bit integer arithmetic. Non-finite values and floats
 * >= 2^64 fall back to snprintf. tests/test_fastfmt.py compares against snprintf over random bit patterns.
 */
#ifndef SIGTK_FASTFMT_H
#define SIGTK_FASTFMT_H
#include <stdint.h>
#include <stdio.h>
#include <string.h>

/* decimal digits of v, returns the end */
static inline char *fmt_u64(char *p, uint64_t v) {
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}
static inline char *fmt_i64(char *p, int64_t v) { /* "%ld" / "%d" */
    if (v < 0) { *p++ = '-'; return fmt_u64(p, (uint64_t)(-(v + 1)) + 1u); }
    return fmt_u64(p, (uint64_t)v);
}
static inline char *fmt_str(char *p, const char *s) {
    const size_t n = strlen(s);
    memcpy(p, s, n);
    return p + n;
}
/* "%f" of a float; writes at most 48 bytes */
static inline char *fmt_f6(char *p, float f) {
    uint32_t b;
    memcpy(&b, &f, 4);
    const uint32_t ex = (b >> 23) & 0xffu, man = b & 0x7fffffu;
    if (ex == 255u || ex >= 150u + 40u) return p + snprintf(p, 48, "%f", (double)f); /* nan, inf, >= 2^64 */
    if (b >> 31) *p++ = '-';
    uint64_t ip, frac;
    if (ex >= 150u) { /* an integer: M * 2^(ex-150) */
        ip = (uint64_t)(man | 0x800000u) << (ex - 150u);
        frac = 0;
    } else {
        const uint64_t num = (uint64_t)(ex ? (man | 0x800000u) : man) * 1000000u; /* < 2^44 */
        const uint32_t s = ex ? 150u - ex : 149u;                                 /* value * 10^6 = num / 2^s */
        uint64_t q = 0;
        if (s < 64u) {
            q = num >> s;
            const uint64_t r = num & ((1ull << s) - 1u), half = 1ull << (s - 1u);
            if (r > half || (r == half && (q & 1u))) q++;
        } /* else num < 2^44 <= 2^(s-1): rounds to 0, a tie is impossible */
        ip = q / 1000000u;
        frac = q % 1000000u;
    }
    p = fmt_u64(p, ip);
    *p++ = '.';
    for (int k = 5; k >= 0; k--) { p[k] = (char)('0' + frac % 10); frac /= 10; }
    return p + 6;
}
#endif
