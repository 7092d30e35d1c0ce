// long_filter_check.cpp -- TEST INFRASTRUCTURE (a checker; never linked into the product).
//
// The fast path does not evaluate the long window's t-statistic (events.c:338-361 with w = w_long); it PROVES, per
// position, that the reference's rounded value cannot exceed the long detector's threshold, with the float test
// long_candidate() of sigtk_b200/csrc/walk_core.cuh on exact integer window sums (derivation: long_filter.md).
// This program checks the implication
//         t2_reference(p) > thr   ==>   long_candidate(...) == true
// on random and adversarial windows, with the VERY code the kernel runs (walk_core.cuh is host/device clean) against
// the oracle's restatement of the reference chain (sigtk_oracle.c: orc_pa, orc_prefix, orc_tstat), and reports how
// tight the test is (how often it fires although t2 <= thr).
//
//   g++ -O2 -std=c++17 -ffp-contract=off long_filter_check.cpp ../sigtk_oracle.c -lm -o /tmp/long_filter_check
//   /tmp/long_filter_check [cases=200000000] [seed=1]
//
// Exit status 1 on any violation.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../sigtk_b200/csrc/walk_core.cuh"
extern "C" {
#include "../sigtk_oracle.h"
}

using namespace sgpu::walk;

static uint64_t rs = 88172645463325252ull;
static inline uint64_t rnd() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return rs; }
static inline double urand() { return (double)(rnd() >> 11) * (1.0 / 9007199254740992.0); }
static inline double nrand() {
    double u1 = urand(), u2 = urand();
    if (u1 < 1e-300) u1 = 1e-300;
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

template <int RNA>
static bool one_case(const int16_t* raw, int c0, double dig, double off, double range, float thr, bool* cand_out, float* t_out) {
    constexpr int w = Cfg<RNA>::w2;
    const int n = 2 * w;
    float pa[2 * 14];
    double S[2 * 14 + 1], Q[2 * 14 + 1];
    float t[2 * 14];
    orc_pa(raw, n, dig, off, range, pa);
    orc_prefix(pa, n, S, Q);
    orc_tstat(S, Q, n, w, t);
    const float tref = t[w];
    // the kernel's side: exact integer sums of z = raw - c0, |z| <= ZMAX
    float SL = 0, QL = 0, SR = 0, QR = 0;
    for (int k = 0; k < w; k++) {
        const float zl = (float)(raw[k] - c0), zr = (float)(raw[w + k] - c0);
        SL += zl; QL += zl * zl; SR += zr; QR += zr * zr;
    }
    const float off_f = (float)off;
    const LongK lk = long_consts<RNA>(c0, off_f, thr);
    const bool cand = long_candidate(SR - SL, SL + SR, QL + QR, lk);   // the kernel's three running sums (exact integers)
    *cand_out = cand;
    *t_out = tref;
    return !(tref > thr) || cand;
}

int main(int argc, char** argv) {
    const long long cases = argc > 1 ? atoll(argv[1]) : 200000000ll;
    rs ^= (uint64_t)(argc > 2 ? atoll(argv[2]) : 1) * 0x9E3779B97F4A7C15ull;
    long long bad = 0, hot = 0, fired_cold = 0, cold = 0, near_cold = 0, near_fired = 0;
    int16_t raw[28];
    for (long long it = 0; it < cases; it++) {
        const int rna = (int)(rnd() & 1);
        const int w = rna ? 14 : 6;
        const int kind = (int)(rnd() % 8);
        // digitisation / range / offset as BLOW5 files hold them, plus odd ones
        double dig = 8192.0, range = 1402.882324, off = (double)(rnd() % 53);
        const int pk = (int)(rnd() % 10);
        if (pk == 0) { range = 748.5801 + 100 * urand(); dig = 2048.0; off = -(double)(rnd() % 300); }
        else if (pk == 1) { off = 10.0 * urand() - 5.0 + 0.25 * (double)(rnd() % 4); }            // fractional offsets
        else if (pk == 2) { off = (double)((int)(rnd() % 4000) - 2000); }
        else if (pk == 3) { range = 50.0 + 3000.0 * urand(); dig = (double)(1 << (9 + rnd() % 6)); }
        const float thr = (rnd() % 4) ? 9.0f : (float)(0.2 + 12.0 * urand());
        // a window pair around a level with noise, a step of a height that puts t near the threshold or anywhere
        const double level = 200.0 + 1500.0 * urand() * ((kind & 1) ? 1.0 : 0.3);
        const double sigma = kind == 7 ? 0.4 + urand() : 1.0 + 30.0 * urand() * urand();
        double step;
        if (kind < 4) {
            // ideal t of a clean step of height h between windows of noise sigma: h / sqrt(2 sigma^2 / w)
            const double target = thr * (0.9 + 0.2 * urand() + ((kind == 3) ? 0.0 : 0.0));
            step = target * sqrt(2.0 * sigma * sigma / w) * (rnd() & 1 ? 1.0 : -1.0);
        } else if (kind < 6) {
            step = (urand() - 0.5) * 800.0;
        } else {
            step = 0.0;
        }
        const int c0 = (int)lrint(level + (urand() - 0.5) * 600.0);
        bool ok_range = true;
        for (int k = 0; k < 2 * w; k++) {
            double v = level + (k >= w ? step : 0.0) + sigma * nrand();
            if (kind == 5) v += (k - w) * (urand() - 0.5) * 4.0;   // ramps
            long q = lrint(v);
            if (q > 32767) q = 32767;
            if (q < -32768) q = -32768;
            raw[k] = (int16_t)q;
            if (abs((int)q - c0) > ZMAX) ok_range = false;
        }
        if (!ok_range) continue;   // such windows are candidates by construction (zdirty)
        bool cand;
        float t;
        const bool fine = rna ? one_case<1>(raw, c0, dig, off, range, thr, &cand, &t)
                              : one_case<0>(raw, c0, dig, off, range, thr, &cand, &t);
        if (!fine) {
            if (bad < 20) {
                fprintf(stderr, "VIOLATION rna=%d thr=%g t=%.9g off=%g dig=%g range=%g c0=%d raw:", rna, thr, t, off, dig, range, c0);
                for (int k = 0; k < 2 * w; k++) fprintf(stderr, " %d", raw[k]);
                fprintf(stderr, "\n");
            }
            bad++;
        }
        if (t > thr) hot++;
        else {
            cold++;
            if (cand) fired_cold++;
            if (t > 0.98f * thr) { near_cold++; if (cand) near_fired++; }
        }
    }
    printf("{\"cases\": %lld, \"t_above_threshold\": %lld, \"violations\": %lld, \"fired_with_t_below_threshold\": %lld, "
           "\"t_below_threshold\": %lld, \"of_which_within_2_percent\": %lld, \"fired_within_2_percent\": %lld}\n",
           cases, hot, bad, fired_cold, cold, near_cold, near_fired);
    return bad ? 1 : 0;
}
