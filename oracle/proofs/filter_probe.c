/* filter_probe.c -- TEST INFRASTRUCTURE (design probe for the filtered t-statistic of round 2).
 *
 * Question: if the walker evaluates an "ideal" t-statistic from the raw integers,
 *      t~ = |D_R - D_L| * sqrt(w) / sqrt(V_L + V_R),   D = sum raw, V = w * sum raw^2 - D^2  (all exact integers),
 * how far is the reference's rounded value (events.c:338-361) from it, relative to the natural scale
 * u * M^2 / cv of the reference's own float roundings, and how often does a detector comparison
 * (events.c:393-431) fall inside that distance?
 *
 * build: gcc -O2 -std=c99 -ffp-contract=off filter_probe.c ../sigtk_oracle.c -lm -o /tmp/filter_probe
 */
#define _POSIX_C_SOURCE 200809L
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../sigtk_oracle.h"

static uint64_t rng_s = 88172645463325252ull;
static double urand(void) {
    rng_s ^= rng_s << 13; rng_s ^= rng_s >> 7; rng_s ^= rng_s << 17;
    return (double)(rng_s >> 11) * (1.0 / 9007199254740992.0);
}
static double nrand(void) {
    double u1 = urand(), u2 = urand();
    if (u1 < 1e-300) u1 = 1e-300;
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

typedef struct { double t; double b; } tb_t;

/* ideal t and its bound at position i for window w; K scales the bound */
static tb_t ideal(const int16_t *raw, double off, int64_t n, int64_t i, int w, double K) {
    tb_t r = {0.0, 0.0};
    if (i < w || i + w > n) return r;
    double DL = 0, EL = 0, DR = 0, ER = 0, M = 0;
    for (int k = 0; k < w; k++) {
        const double a = raw[i - w + k] + off, b = raw[i + k] + off;
        DL += a; EL += a * a; DR += b; ER += b * b;
        if (fabs(a) > M) M = fabs(a);
        if (fabs(b) > M) M = fabs(b);
    }
    const double VL = w * EL - DL * DL, VR = w * ER - DR * DR, V = VL + VR;
    const double u = ldexp(1.0, -24);
    if (V <= 0) { r.t = 0; r.b = INFINITY; return r; }
    r.t = fabs(DR - DL) * sqrt((double)w) / sqrt(V);
    /* cv noise: K u M^2 (in raw units: cv = V / w^2);  delta noise: 4 u M */
    const double rel_cv = K * u * M * M * (double)w * w / V;
    if (rel_cv >= 0.5) { r.b = INFINITY; return r; }
    r.b = r.t * (1.0 / sqrt(1.0 - rel_cv) - 1.0) + 4.0 * u * M * sqrt((double)w) * (double)w / sqrt(V * (1.0 - rel_cv)) + 1e-6 * r.t;
    return r;
}

typedef struct { int64_t pp; int valid; double pv, pb; int64_t mt; } idet_t;  /* pv: approximate value, pb: its bound */

int main(int argc, char **argv) {
    const int rna = argc > 1 ? atoi(argv[1]) : 0;
    const int n_reads = argc > 2 ? atoi(argv[2]) : 20;
    const int64_t n = argc > 3 ? atoll(argv[3]) : 40000;
    const double K = argc > 4 ? atof(argv[4]) : 10.0;
    const double noise = argc > 5 ? atof(argv[5]) : 2.0;
    orc_params_t P;
    orc_params(rna, &P);
    const double dig = 8192.0, range = 1402.882324;
    int16_t *raw = malloc(n * sizeof(int16_t));
    float *pa = malloc(n * sizeof(float));
    double *S = malloc((n + 1) * sizeof(double)), *Q = malloc((n + 1) * sizeof(double));
    float *t[2] = {malloc(n * sizeof(float)), malloc(n * sizeof(float))};
    tb_t *id[2] = {malloc(n * sizeof(tb_t)), malloc(n * sizeof(tb_t))};
    double maxratio[2] = {0, 0};
    uint64_t npos = 0, outside[2] = {0, 0};
    uint64_t steps = 0, amb_steps = 0, ambA = 0, ambB = 0, ambC = 0, ambD = 0, ambE = 0, infb = 0, amb_noA = 0;
    double hist_ratio[2][12];
    memset(hist_ratio, 0, sizeof hist_ratio);
    for (int r = 0; r < n_reads; r++) {
        const double off = (double)(r % 53);
        double level = 60 + 60 * urand();
        const double pch = rna ? 0.025 : 0.1;
        for (int64_t i = 0; i < n; i++) {
            if (urand() < pch) level = 60 + 60 * urand();
            double v = rint((level + noise * nrand()) * (dig / range) - off);
            if (v > 32767) v = 32767;
            if (v < -32768) v = -32768;
            raw[i] = (int16_t)v;
        }
        orc_pa(raw, n, dig, off, range, pa);
        orc_prefix(pa, n, S, Q);
        orc_tstat(S, Q, n, P.w_short, t[0]);
        orc_tstat(S, Q, n, P.w_long, t[1]);
        for (int k = 0; k < 2; k++) {
            const int w = k ? P.w_long : P.w_short;
            for (int64_t i = 0; i < n; i++) {
                id[k][i] = ideal(raw, off, n, i, w, K);
                if (i >= w && i + w <= n) {
                    if (isinf(id[k][i].b)) { infb++; continue; }
                    const double ratio = fabs((double)t[k][i] - id[k][i].t) / id[k][i].b;
                    if (ratio > maxratio[k]) maxratio[k] = ratio;
                    if (ratio > 1.0) outside[k]++;
                    int bin = (int)(ratio * 10.0);
                    if (bin > 11) bin = 11;
                    hist_ratio[k][bin]++;
                }
            }
            npos += n;
        }
        /* the true detector on the reference's t, with the interval test on the ideal values next to it */
        idet_t d[2];
        for (int k = 0; k < 2; k++) { d[k].pp = -1; d[k].valid = 0; d[k].pv = FLT_MAX; d[k].pb = 0; d[k].mt = 0; }
        float tpv[2] = {FLT_MAX, FLT_MAX};  /* true peak values */
        const float thr[2] = {P.thr_short, P.thr_long};
        const int win[2] = {(int)P.w_short, (int)P.w_long};
        const float h = P.peak_height;
        for (int64_t i = 0; i < n; i++) {
            int amb = 0, amb_wo_a = 0;
            for (int k = 0; k < 2; k++) {
                idet_t *D = &d[k];
                if (D->mt >= i) continue;
                const float cur = t[k][i];
                const double c = id[k][i].t, cb = id[k][i].b;
                steps++;
                if (D->pp < 0) {
                    /* A: cur < pv ; B: cur - pv > h */
                    if (tpv[k] != FLT_MAX) {
                        if (fabs(c - D->pv) <= cb + D->pb) { ambA++; amb = 1; }
                        if (fabs(c - D->pv - h) <= cb + D->pb) { ambB++; amb = 1; amb_wo_a = 1; }
                    }
                    if (cur < tpv[k]) { tpv[k] = cur; D->pv = c; D->pb = cb; }
                    else if (cur - tpv[k] > h) { tpv[k] = cur; D->pv = c; D->pb = cb; D->pp = i; }
                    continue;
                }
                if (fabs(c - D->pv) <= cb + D->pb) { ambC++; amb = 1; amb_wo_a = 1; }
                if (cur > tpv[k]) { tpv[k] = cur; D->pv = c; D->pb = cb; D->pp = i; }
                if (fabs(D->pv - thr[k]) <= D->pb) { ambD++; amb = 1; amb_wo_a = 1; }
                if (k == 0 && tpv[0] > thr[0]) {
                    d[1].mt = D->pp + win[0]; d[1].pp = -1; d[1].valid = 0; tpv[1] = FLT_MAX; d[1].pv = FLT_MAX; d[1].pb = 0;
                }
                if (tpv[k] > thr[k] && !D->valid && fabs(D->pv - c - h) <= cb + D->pb) { ambE++; amb = 1; amb_wo_a = 1; }
                if (tpv[k] - cur > h && tpv[k] > thr[k]) D->valid = 1;
                if (D->valid && (i - D->pp) > win[k] / 2) {
                    D->pp = -1; tpv[k] = cur; D->pv = c; D->pb = cb; D->valid = 0;
                }
            }
            amb_steps += amb;
            amb_noA += amb_wo_a;
        }
    }
    const double ns = (double)n_reads * (double)n;
    printf("rna=%d reads=%d n=%ld K=%.1f noise=%.2f\n", rna, n_reads, (long)n, K, noise);
    for (int k = 0; k < 2; k++) {
        printf("  window %d: max |t_ref - t_ideal| / bound = %.4f   outside=%lu\n    hist(ratio*10):", k, maxratio[k], (unsigned long)outside[k]);
        for (int b = 0; b < 12; b++) printf(" %.0f", hist_ratio[k][b]);
        printf("\n");
    }
    printf("  infinite bounds: %lu\n", (unsigned long)infb);
    printf("  ambiguous sample-steps: %.3e per sample (without A: %.3e)  [A %.2e B %.2e C %.2e D %.2e E %.2e per sample]\n",
           amb_steps / ns, amb_noA / ns, ambA / ns, ambB / ns, ambC / ns, ambD / ns, ambE / ns);
    (void)steps; (void)npos;
    return 0;
}
