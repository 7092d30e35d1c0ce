// walk_core.cuh -- the per-thread CHUNK WALKER of the fast path (walk.cu).
//
// One thread walks one chunk of one read sample by sample, entirely in registers:
//
//   int16 -> pA (misc.c:26-29) -> running FP64 prefix of x and x*x (a small register ring) -> the window sums of
//   both window lengths (w2 = 2*w1: the long window is the sum of two short ones) -> both t-statistics
//   (events.c:338-361, every rounding step written out) -> both peak detectors (events.c:387-437).
//
// Nothing per-sample is staged in shared or global memory: the only per-sample traffic is the 2-byte load (and
// the optional 4-byte pA store); emitted peaks are collected in a per-block register mask and set in the event-start
// bitmap with fire-and-forget RED.OR.
//
// Timeline of one walker step (newest sample index j, all indices relative to the read):
//     P(j+1)            = P(j) + x[j]                      running prefix sums (exact, see below)
//     D1(j1), j1=j-w1+1 = P(j+1) - P(j1)                   short-window sums, ring of w1
//     D2(j2), j2=j1-w1  = D1(j2) + D1(j1)                  long-window sums
//     t1(j1)  from the window terms of D1(j1-w1), D1(j1)   -> ring of w1
//     t2(j2)  from the window terms of D2(j2-w2), D2(j2)
//     detector step at position p = j2 = j - (w2-1) with t1(p) (from the ring) and t2(p)
//
// Why any summation order gives the reference's bits: the reference's S[i], Q[i] are sequential double sums of
// floats. If every value of a read is a multiple of 2^k and the sum of magnitudes stays below 2^(k+53), no
// addition ever rounds, every partial sum of any subset is exact, and S[b]-S[a] is the exact sum of the samples
// in [a,b) however it is formed. The per-read witness (min/max |pA| collected by the walker, evaluated by
// build_seq_list_kernel) checks that sufficient condition; reads that fail are redone by the sequential-order
// kernels (generic.cu).
//
// Rare paths, all inside the walker (the read stays on the fast path): a t-statistic next to a float rounding
// midpoint or a peak older than the block's mask -> redo_block (the block's t-statistics from the raw samples with
// the reference's own operations); LOW samples (pA <= 0 or barely above: glitches) -> handed to the running sums as
// 0, the three blocks around them redone (see walk_block).
//
// The file is `__host__ __device__` clean so that tests/tools/host_walk.cpp can run the very same chunk logic on
// the CPU against the oracle (a checker for the chunking / ring / boundary logic; never part of the product).
#pragma once
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define SGW_HD __host__ __device__ __forceinline__
#else
#define SGW_HD inline
#endif

namespace sgpu {
namespace walk {

// ---- arithmetic with explicit rounding (device: intrinsics, never contracted; host: plain IEEE ops) -------------
#if defined(__CUDA_ARCH__)
SGW_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
SGW_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
SGW_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
SGW_HD double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }
SGW_HD double ddiv(double a, double b) { return __ddiv_rn(a, b); }
SGW_HD double dsqrt(double a) { return __dsqrt_rn(a); }
SGW_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
SGW_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
SGW_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
SGW_HD float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
SGW_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
SGW_HD float d2f(double a) { return __double2float_rn(a); }
SGW_HD uint32_t d_hi(double a) { return (uint32_t)__double2hiint(a); }
SGW_HD uint32_t d_lo(double a) { return (uint32_t)__double2loint(a); }
SGW_HD uint32_t f_bits(float a) { return __float_as_uint(a); }
SGW_HD float bits_f(uint32_t b) { return __uint_as_float(b); }
SGW_HD float rsqrt_seed(float a) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a));
    return y;
}
// (double)a for a POSITIVE NORMAL float, exactly, with one integer multiply-add on the FMA pipe instead of a
// conversion on the (16 lanes/clk/SM) XU pipe: double bits = float bits * 2^29 + (1023-127) * 2^52.
// Any other input gives garbage; callers only use the result where the input is known to be positive and normal.
SGW_HD double widen_pos(float a) {
    return __longlong_as_double((long long)((unsigned long long)__float_as_uint(a) * 0x20000000ull + 0x3800000000000000ull));
}
SGW_HD double widen_abs(float a) { return (double)fabsf(a); }  // any a (one conversion; 0 stays 0)
SGW_HD uint32_t shr_clamp(uint32_t x, int s) {  // x >> s; 0 when s is outside [0, 31] (PTX shr clamps the amount)
    uint32_t r;
    asm("shr.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(s));
    return r;
}
// Two independent IEEE float operations in ONE instruction (FMUL2 / FFMA2 / FADD2 of sm_100): each half is
// rounded exactly like the scalar operation, so pairing never changes a bit; it halves the issue slots of the
// float side of the window terms.
struct F2 { float lo, hi; };
SGW_HD F2 f2mul(F2 a, F2 b) {
    F2 r;
    asm("{\n\t.reg .b64 a, b, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tmul.rn.f32x2 d, a, b;\n\t"
        "mov.b64 {%0, %1}, d;\n\t}" : "=f"(r.lo), "=f"(r.hi) : "f"(a.lo), "f"(a.hi), "f"(b.lo), "f"(b.hi));
    return r;
}
SGW_HD F2 f2sq(F2 a) {  // {a.lo * a.lo, a.hi * a.hi}
    F2 r;
    asm("{\n\t.reg .b64 a, d;\n\tmov.b64 a, {%2, %3};\n\tmul.rn.f32x2 d, a, a;\n\tmov.b64 {%0, %1}, d;\n\t}"
        : "=f"(r.lo), "=f"(r.hi) : "f"(a.lo), "f"(a.hi));
    return r;
}
SGW_HD F2 f2add(F2 a, F2 b) {
    F2 r;
    asm("{\n\t.reg .b64 a, b, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tadd.rn.f32x2 d, a, b;\n\t"
        "mov.b64 {%0, %1}, d;\n\t}" : "=f"(r.lo), "=f"(r.hi) : "f"(a.lo), "f"(a.hi), "f"(b.lo), "f"(b.hi));
    return r;
}
SGW_HD F2 f2sub(F2 a, F2 b) {
    F2 r;
    asm("{\n\t.reg .b64 a, b, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tsub.rn.f32x2 d, a, b;\n\t"
        "mov.b64 {%0, %1}, d;\n\t}" : "=f"(r.lo), "=f"(r.hi) : "f"(a.lo), "f"(a.hi), "f"(b.lo), "f"(b.hi));
    return r;
}
SGW_HD F2 f2fma(F2 a, F2 b, F2 c) {
    F2 r;
    asm("{\n\t.reg .b64 a, b, c, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tmov.b64 c, {%6, %7};\n\t"
        "fma.rn.f32x2 d, a, b, c;\n\tmov.b64 {%0, %1}, d;\n\t}"
        : "=f"(r.lo), "=f"(r.hi) : "f"(a.lo), "f"(a.hi), "f"(b.lo), "f"(b.hi), "f"(c.lo), "f"(c.hi));
    return r;
}
#else
SGW_HD double dadd(double a, double b) { return a + b; }
SGW_HD double dsub(double a, double b) { return a - b; }
SGW_HD double dmul(double a, double b) { return a * b; }
SGW_HD double dfma(double a, double b, double c) { return fma(a, b, c); }
SGW_HD double ddiv(double a, double b) { return a / b; }
SGW_HD double dsqrt(double a) { return sqrt(a); }
SGW_HD float fadd(float a, float b) { return a + b; }
SGW_HD float fsub(float a, float b) { return a - b; }
SGW_HD float fmul(float a, float b) { return a * b; }
SGW_HD float ffma(float a, float b, float c) { return fmaf(a, b, c); }
SGW_HD float fdiv(float a, float b) { return a / b; }
SGW_HD float d2f(double a) { return (float)a; }
SGW_HD uint32_t d_hi(double a) { uint64_t b; memcpy(&b, &a, 8); return (uint32_t)(b >> 32); }
SGW_HD uint32_t d_lo(double a) { uint64_t b; memcpy(&b, &a, 8); return (uint32_t)b; }
SGW_HD uint32_t f_bits(float a) { uint32_t b; memcpy(&b, &a, 4); return b; }
SGW_HD float bits_f(uint32_t b) { float a; memcpy(&a, &b, 4); return a; }
SGW_HD float rsqrt_seed(float a) { return (float)(1.0 / sqrt((double)a)); }  // the guard below absorbs +-4 ulp
SGW_HD double widen_pos(float a) {  // same bit arithmetic as the device version (garbage for non-positive / denormal a)
    uint32_t b; memcpy(&b, &a, 4);
    const uint64_t v = (uint64_t)b * 0x20000000ull + 0x3800000000000000ull;
    double r; memcpy(&r, &v, 8); return r;
}
SGW_HD double widen_abs(float a) { return (double)fabsf(a); }
SGW_HD uint32_t shr_clamp(uint32_t x, int s) { return (unsigned)s < 32u ? x >> s : 0u; }
struct F2 { float lo, hi; };
SGW_HD F2 f2mul(F2 a, F2 b) { F2 r; r.lo = a.lo * b.lo; r.hi = a.hi * b.hi; return r; }
SGW_HD F2 f2sq(F2 a) { F2 r; r.lo = a.lo * a.lo; r.hi = a.hi * a.hi; return r; }
SGW_HD F2 f2add(F2 a, F2 b) { F2 r; r.lo = a.lo + b.lo; r.hi = a.hi + b.hi; return r; }
SGW_HD F2 f2sub(F2 a, F2 b) { F2 r; r.lo = a.lo - b.lo; r.hi = a.hi - b.hi; return r; }
SGW_HD F2 f2fma(F2 a, F2 b, F2 c) { F2 r; r.lo = fmaf(a.lo, b.lo, c.lo); r.hi = fmaf(a.hi, b.hi, c.hi); return r; }
#endif

// ---- parameters (events.c:35-54) ----------------------------------------------------------------------------------
template <int RNA>
struct Cfg {
    static constexpr int w1 = RNA ? 7 : 3;
    static constexpr int w2 = 2 * w1;
    static constexpr int LAG = w2 - 1;        // detector position = newest sample index - LAG
    static constexpr int R1 = RNA ? 8 : 4;    // ring size of the depth-w1 rings (> w1, power of two)
    static constexpr int R2 = RNA ? 16 : 8;   // ring size of the depth-w2 rings (> w2, power of two)
    static constexpr int U = RNA ? 16 : 8;    // samples per block (multiple of R2: every ring index is static;
                                              // 16 for DNA measured 24 % slower: registers, and a shorter peak mask)
    static constexpr int FILL = 2;            // blocks that only fill the rings before the first detector step
    static_assert(FILL * U >= 2 * w2 - 1, "fill covers the look-back of the first step");
};
template <int RNA> SGW_HD float thr_short() { return RNA ? 2.5f : 1.4f; }
template <int RNA> SGW_HD float thr_long() { return 9.0f; }
template <int RNA> SGW_HD float peak_h() { return RNA ? 1.0f : 0.2f; }

// ---- exact shortcuts (validated on the CPU by oracle/proofs/*.c) --------------------------------------------------
// a / W, W in {3,6,7,14}: reciprocal multiply + one FMA residual correction (Markstein). Equal to the IEEE
// quotient for every double the path can produce, and for every float with |a| >= 2^-125 or a == +0
// (exhaustive check over all floats). The walker only uses the float form on values that the per-read witness
// certifies to be in that range (nonzero |pA| >= 2^-60, unit > 0), or on cv >= 2e-29.
template <int W>
SGW_HD double ddivw(double a) {
    constexpr double r = 1.0 / (double)W;
    const double q0 = dmul(a, r);
    const double e = dfma(-(double)W, q0, a);
    return dfma(e, r, q0);
}
template <int W>
SGW_HD float fdivw(float a) {
    constexpr float r = 1.0f / (float)W;
    const float q0 = fmul(a, r);
    const float e = ffma(-(float)W, q0, a);
    return ffma(e, r, q0);
}

template <int WA, int WB>
SGW_HD F2 fdivw2(F2 a) {  // {a.lo / WA, a.hi / WB}: fdivw on both halves at once
    const F2 r = {1.0f / (float)WA, 1.0f / (float)WB};
    const F2 nw = {-(float)WA, -(float)WB};
    const F2 q0 = f2mul(a, r);
    const F2 e = f2fma(nw, q0, a);
    return f2fma(e, r, q0);
}

// t = (float)(fabs((double)delta) / sqrt((double)scaled)), scaled = cv / W  (events.c:360) through the 22-bit
// reciprocal square root y0 of the hardware and ONE second-order correction in double:
//     eh = 1/2 - (scaled/2) * y0^2,   q = |delta| * y0 * (1 + eh)
// With y0 = (1 + eps) / sqrt(scaled), |eps| <= 2^-22.4 (PTX rsqrt.approx), the method error of q is 1.5 eps^2 <=
// 2^-44.2 relative = at most 445 ulp(double), plus 2 ulp of rounding. The float rounding of q equals the
// reference's doubly rounded value unless q lies within that distance of a float rounding midpoint
// (oracle/proofs/tstat_tail_check.c); values within 1024 ulp(double) of a midpoint (about 4 in a million) and
// cv < 2e-29 clear `ok`, and the caller recomputes the block with the reference's own operations (redo_block).
// No range check is needed on q: the per-read witness bounds every pA to [2^-60, 2^20), so |delta| is 0 or in
// [2^-84, 2^21], cv <= 2^42 and (checked here) cv >= 2e-29, hence q == 0 or 2^-105 < q < 2^72: a normal float
// after rounding. delta == 0 gives q == +0 and t == +0 like the reference (flat stretches).
SGW_HD float tail(float delta, float scaled, float scaled_half, float cv, bool& ok) {
    const double ch = widen_pos(scaled_half);              // scaled / 2 (exact); scaled >= 1e-30 when cv >= 2e-29
    const double y0 = widen_pos(rsqrt_seed(scaled));
    const double th = dmul(ch, y0);
    const double eh = dfma(-th, y0, 0.5);
    const double q0 = dmul(widen_abs(delta), y0);
    const double q = dfma(q0, eh, q0);
    // accept when the 29 bits below float precision are not within 1024 of the midpoint (low word, shifted up)
    ok = ok & ((d_lo(q) * 8u - ((0x10000000u - 1024u) << 3)) >= (2048u << 3)) & (cv >= 2.0e-29f);
    return d2f(q);
}

// The terms of events.c:339-350 that depend on ONE window [j, j+W) only. The window is the right window of
// position j (sum2/sumsq2, narrowed to float at once) and the left window of position j+W (kept in double).
// The float products and quotients widened here are positive and normal for every window that lies inside a
// read that passed the witness (all pA > 0, >= 2^-60).
template <int W>
SGW_HD void window_terms(double D, double E, float& A, double& Lq, float& B, double& Vd, double& B2d) {
    A = d2f(ddivw<W>(D));                                 // mean1 = (float)(sum1 / w)
    const F2 de = {d2f(D), d2f(E)};
    const F2 bv = fdivw2<W, W>(de);                       // mean2 = (float)sum2 / w ; (float)sumsq2 / w
    Lq = dsub(ddivw<W>(E), widen_pos(fmul(A, A)));        // sumsq1/w - (double)(mean1*mean1)
    B = bv.lo;
    Vd = widen_pos(bv.hi);
    B2d = widen_pos(fmul(B, B));                          // (double)(mean2*mean2)
}
// the short-window t-statistics of TWO consecutive positions: the windows' terms -> combined variance
// (events.c:349-353) -> tail; the float divisions and halvings of the two positions share packed instructions.
// No fmaxf(cv, FLT_MIN) (events.c:353) here: `tail` clears ok for every cv below 2e-29 (NaN included), and such a
// block is recomputed with the reference's own operations, floor included.
template <int W>
SGW_HD void tstat_two(float A0, double L0, float B0, double V0, double B0sq, float A1, double L1, float B1, double V1,
                      double B1sq, float& t0, bool& ok0, float& t1, bool& ok1) {
    const double acc0 = dsub(dadd(L0, V0), B0sq);         // ((.. - m1sq) + v2) - m2sq, left to right in double
    const double acc1 = dsub(dadd(L1, V1), B1sq);
    const F2 cv = {d2f(acc0), d2f(acc1)};
    const F2 sc = fdivw2<W, W>(cv);                       // cv / w (float); shortcut valid for cv >= 1e-36
    const F2 half = {0.5f, 0.5f};
    const F2 sh = f2mul(sc, half);
#if defined(WALK_X_NOTAIL)
    t0 = fmul(fsub(B0, A0), sc.lo); t1 = fmul(fsub(B1, A1), sh.hi);
#else
    t0 = tail(fsub(B0, A0), sc.lo, sh.lo, cv.lo, ok0);    // delta = mean2 - mean1
    t1 = tail(fsub(B1, A1), sc.hi, sh.hi, cv.hi, ok1);
#endif
}

// ---- the reference's own operation sequence, from the raw samples (rare path) ---------------------------------------
template <class Io>
SGW_HD float sample_pa(const Io& io, int i, int n, float off, float unit) {
    if (i < 0 || i >= n) return 0.0f;
    int v[4];
    io.load8(i & ~7, v);
    const int q = i & 7;
    const int word = q < 2 ? v[0] : q < 4 ? v[1] : q < 6 ? v[2] : v[3];
    const int raw = (q & 1) ? (word >> 16) : (int)(int16_t)(word & 0xffff);
    return fmul(fadd((float)raw, off), unit);
}
template <class Io>
SGW_HD void window_sums(const Io& io, int a, int w, int n, float off, float unit, double& S, double& Q) {
    S = 0.0; Q = 0.0;
    for (int i = a; i < a + w; i++) {
        const float x = sample_pa(io, i, n, off, unit);
        S = dadd(S, (double)x);
        Q = dadd(Q, (double)fmul(x, x));
    }
}
// events.c:338-361 from the four window sums of one position (sum1 / ssq1: the left window, kept in double;
// sum2d / ssq2d: the right window, narrowed to float at once), every operation the reference's own
SGW_HD float tstat_from_sums(double sum1, double ssq1, double sum2d, double ssq2d, int w) {
    const float wf = (float)w;
    const float sum2 = d2f(sum2d), ssq2 = d2f(ssq2d);
    const float mean1 = d2f(ddiv(sum1, (double)wf));
    const float mean2 = fdiv(sum2, wf);
    double acc = ddiv(ssq1, (double)wf);
    acc = dsub(acc, (double)fmul(mean1, mean1));
    acc = dadd(acc, (double)fdiv(ssq2, wf));
    acc = dsub(acc, (double)fmul(mean2, mean2));
    const float cv = fmaxf(d2f(acc), FLT_MIN);
    const float delta = fsub(mean2, mean1);
    return d2f(ddiv(fabs((double)delta), dsqrt((double)fdiv(cv, wf))));
}
// the same for position i with window length w from the raw samples (the sums are exact, so forming them per
// window is the same as the reference's prefix differences)
template <class Io>
SGW_HD float tstat_exact(const Io& io, int i, int w, int n, float off, float unit) {
    double sum1, ssq1, sum2d, ssq2d;
    window_sums(io, i - w, w, n, off, unit, sum1, ssq1);
    window_sums(io, i, w, n, off, unit, sum2d, ssq2d);
    return tstat_from_sums(sum1, ssq1, sum2d, ssq2d, w);
}
// t-statistics of CONSECUTIVE positions: the four window sums slide by one sample per step (three samples touched
// instead of 2w; exact like every other grouping of these sums). next(i) must be called with i increasing by one;
// positions whose windows leave the read give 0 (events.c:328-338).
template <class Io>
struct SlidingT {
    double s1, q1, s2, q2;
    int at;   // the position the sums belong to, -1: none
    SGW_HD SlidingT() : s1(0.0), q1(0.0), s2(0.0), q2(0.0), at(-1) {}
    SGW_HD float next(const Io& io, int i, int w, int n, float off, float unit) {
        if (!(i >= w && i + w <= n)) { at = -1; return 0.0f; }
        if (at != i - 1) {
            window_sums(io, i - w, w, n, off, unit, s1, q1);
            window_sums(io, i, w, n, off, unit, s2, q2);
        } else {
            const float xa = sample_pa(io, i - 1 - w, n, off, unit);   // leaves the left window
            const float xb = sample_pa(io, i - 1, n, off, unit);       // moves from the right window to the left one
            const float xc = sample_pa(io, i - 1 + w, n, off, unit);   // enters the right window
            const double da = (double)xa, db = (double)xb, dc = (double)xc;
            const double qa = (double)fmul(xa, xa), qb = (double)fmul(xb, xb), qc = (double)fmul(xc, xc);
            s1 = dadd(dsub(s1, da), db); q1 = dadd(dsub(q1, qa), qb);
            s2 = dadd(dsub(s2, db), dc); q2 = dadd(dsub(q2, qb), qc);
        }
        at = i;
        return tstat_from_sums(s1, q1, s2, q2, w);
    }
};
// ---- the peak detectors (events.c:371-443) ------------------------------------------------------------------------
// Positions are "shifted" read indices: index in the read + (read_off & 31), so that position >> 5 is a word of the
// read's part of the event-start bitmap; they stay below 2^30 (reads of 2^30 samples or more take the
// sequential-order kernels). One word `ps` per detector packs peak_pos and valid_peak of the reference's Detector
// (events.c:269-281):   ps = PS_NONE              peak_pos == -1 (CASE 1 of events.c:393)
//                       ps = peak_pos | PS_OPEN   CASE 2, valid_peak == false
//                       ps = peak_pos             CASE 2, valid_peak == true
// so that "valid_peak && i - peak_pos > w/2" (events.c:429) is the single comparison  i - ps > w/2.
//
// THE LONG DETECTOR IS NOT STEPPED ON THE FAST PATH (round 2). The short detector is autonomous; the long one is
// driven by it: every step at which the short detector holds a peak above its threshold resets the long detector
// and masks it up to peak_pos + w_short (events.c:414-422). Call the steps between two such resets a LIFE of the
// long detector. Within a life its peak_value is FLT_MAX (CASE 1, before its first step) or the t-statistic of
// one of the life's stepped positions, and valid_peak needs peak_value > threshold (events.c:424): a life in which
// no stepped position has t2 > thr_long emits nothing, and the state it ends in is wiped by the next reset. On
// real R9.4 reads and on the synthetic sets the long detector never emits at all (0 of 92,935 events of
// sp1_dna.blow5) and about 2 steps in 10,000 are stepped with t2 > 9.
//   The walker therefore only tracks, per step, where the current life's steps start (`l_start`: a function of the
// short detector alone) and whether a stepped position of the life MAY have t2 > thr_long -- decided by a
// conservative float test on exact integer window sums of the raw samples (long_candidate below; the bound is
// proved in oracle/proofs/long_filter.md and checked by oracle/proofs/long_filter_check.c). A life that may be
// hot becomes a JOB: long_job() replays the long detector over exactly that life with the reference's own
// operations (tstat_exact) and records whatever it emits (long_jobs_kernel in walk.cu; OR into the bitmap is
// idempotent). Everything else about the long window -- its t-statistic chain and its detector -- is gone from
// the per-sample path.
constexpr int PS_NONE = 0x7fff0000;  // (room below INT_MAX: position - PS_NONE must not wrap for positions >= -2^16)
constexpr int PS_OPEN = 0x40000000;
constexpr int LS_PRED = -0x40000000;  // l_start: the life began before this chunk's first owned step (see long_job)
constexpr int LS_CONT = 0x7fffffff;   // job end: the life runs past the chunk's last owned step
struct WalkDet {
    float s_pv; int s_ps;               // short detector (never masked after the read's first sample)
    int l_start;                        // first step of the long detector's current life that is not masked:
                                        // max(reset step, masked_to + 1); LS_PRED while unknown to this chunk
    bool l_hot;                         // a stepped position of the current life may have t2 > thr_long
};
SGW_HD void det_cold(WalkDet& d, int first_step) {  // both detectors start at `first_step` from the reset state
    d.s_pv = FLT_MAX; d.s_ps = PS_NONE;
    d.l_start = first_step; d.l_hot = false;
}

// canonical form at boundary b (state before step b): the first 6 words are bit-compared between the chunk that
// ends at b and the chunk that starts there (the short detector; the long one has no state on the fast path);
// word 6 of an END record carries l_start for the jobs of later chunks (not compared).
struct Canon { int v[8]; };
template <int RNA>
SGW_HD Canon canon_of(const WalkDet& d, int b) {
    Canon c;
    (void)b;
    c.v[0] = (int)f_bits(d.s_pv); c.v[1] = d.s_ps; c.v[2] = 0; c.v[3] = 0; c.v[4] = 0; c.v[5] = 0;
    c.v[6] = d.l_start; c.v[7] = 0;
    return c;
}

// Emitted peaks of one block are collected in a register: bit k of `mk` <=> a peak at position ub + k. The
// block's own steps sit at bits PK_LEAD .. PK_LEAD + U - 1, so a peak up to PK_LEAD positions older than the
// step that emits it fits; an older one (a plateau within peak_height of a maximum above the threshold lasting
// dozens of samples; seen in about 1 RNA read in 20,000) shows in `oldest`, and the block is redone by redo_block,
// which records every peak on its own.
template <int RNA> struct PkCfg {
    static constexpr int LEAD = 32 - Cfg<RNA>::U;
#if defined(WALK_TEST_FAR)  // host tests: treat every peak older than w/2 + 2 as too old for the mask
    static constexpr int FAR = 0;
#else
    static constexpr int FAR = LEAD - (Cfg<RNA>::w1 / 2 + 1);  // largest (age - w/2 - 1) that always fits the mask
#endif
};
// (`jobs`, `job_ls`, `job_end`: the fast path does not call io.job() from inside a block, because the block may
// still be redone with exact values; it parks the block's first job here and a second one sends the block to redo_block)
struct PeakAcc { uint32_t mk; int oldest; int jobs, job_ls, job_end; };
struct NoEmit { SGW_HD void operator()(int) const {} };

// One detector, one position (events.c:393-437); the detector is not masked at u (events.c:387).
//   m    : index of the step within its block (compile-time), u : its position
//   c    : the t-statistic at u, thr : the detector's threshold
//   big2 : CASE 2 and the running maximum is above the threshold (the short detector then masks the long one)
//   p2   : the (possibly moved) peak position with the PS_OPEN flag, valid when big2
//   on_emit(pos) : called for an emitted peak (the fast path passes NoEmit and reads the mask instead)
template <bool SHORT, int RNA, class E>
SGW_HD int det_one(float& pv, int& ps, int m, int u, float c, float thr, PeakAcc& acc, bool& big2, int& p2,
                   const E& on_emit) {
    constexpr int w = SHORT ? Cfg<RNA>::w1 : Cfg<RNA>::w2;
    const float h = peak_h<RNA>();
    const bool none = ps == PS_NONE;
    const float df = fsub(c, pv);                          // > 0: above the running value, < 0: below
    const float pvm = none ? fminf(pv, c) : fmaxf(pv, c);  // CASE 1: running minimum (394); CASE 2: running maximum (408)
    const bool rise1 = none & (df > h);                    // CASE 1 -> CASE 2 at u (399-403)
    const bool gt2 = !none & (df > 0.0f);                  // CASE 2: a new maximum moves the peak, valid_peak is kept
    big2 = !none & (pvm > thr);
    const bool drop = big2 & (df < -h);                    // pv - c > h (the negation is exact) && pv > threshold (424)
    p2 = gt2 ? ((ps & PS_OPEN) | u) : ps;
    const int p3 = drop ? (p2 & ~PS_OPEN) : p2;            // valid_peak = true
    const int over = (u - (w / 2 + 1)) - p3;               // age of the peak beyond w/2 + 1; negative in CASE 1 and
    const bool emit = over >= 0;                           // while not valid (flag bits): events.c:429
    acc.mk |= shr_clamp(1u << (PkCfg<RNA>::LEAD + m - (w / 2 + 1)), over);  // no bit unless 0 <= over <= 31
    if (emit) on_emit(p3);
    pv = (rise1 | emit) ? c : pvm;                         // 395 / 400 / 409 / 433
    ps = emit ? PS_NONE : p3;
    ps = rise1 ? (u | PS_OPEN) : ps;
    return over;
}

// Both detectors stepped together, short first (events.c:385-440): the sequential-order kernels (generic.cu) walk
// stored t arrays with this; the fast path does not step the long detector (see above).
struct DualDet {
    float s_pv; int s_ps;               // short detector
    float l_pv; int l_ps; int l_mt;     // long detector; l_mt = masked_to
};
SGW_HD void dual_cold(DualDet& d, int first_step) {
    d.s_pv = FLT_MAX; d.s_ps = PS_NONE;
    d.l_pv = FLT_MAX; d.l_ps = PS_NONE; d.l_mt = first_step - 1;
}
SGW_HD Canon dual_canon(const DualDet& d, int b) {
    Canon c;
    c.v[0] = (int)f_bits(d.s_pv); c.v[1] = d.s_ps; c.v[2] = (int)f_bits(d.l_pv); c.v[3] = d.l_ps;
    c.v[4] = d.l_mt >= b ? d.l_mt : -1;
    c.v[5] = 0; c.v[6] = 0; c.v[7] = 0;
    return c;
}
template <int RNA, class E>
SGW_HD void dual_step(DualDet& d, int u, float c1, float c2, const E& on_emit) {
    bool maskl, unused_b; int p2, unused_p;
    PeakAcc unused; unused.mk = 0u; unused.oldest = 0;
    det_one<true, RNA>(d.s_pv, d.s_ps, 0, u, c1, thr_short<RNA>(), unused, maskl, p2, on_emit);
    // the short detector dominates the long one while it holds a peak above its threshold (events.c:414-422)
    d.l_mt = maskl ? (p2 & ~PS_OPEN) + Cfg<RNA>::w1 : d.l_mt;
    d.l_ps = maskl ? PS_NONE : d.l_ps;
    d.l_pv = maskl ? FLT_MAX : d.l_pv;
    // a masked long detector is always in the reset state (masked_to is only ever set together with a reset, and
    // a masked detector is not stepped), and the reset state does not react to FLT_MAX: no gating needed
    const float c2m = d.l_mt < u ? c2 : FLT_MAX;
    det_one<false, RNA>(d.l_pv, d.l_ps, 0, u, c2m, thr_long<RNA>(), unused, unused_b, unused_p, on_emit);
}

// ---- the long window on the fast path: a conservative test on exact integer sums -----------------------------------
// z = raw - c0 (c0: an integer pivot of the chunk, |z| <= ZMAX so that every sum below is an exact integer in
// float). For the windows L = [p - w, p), R = [p, p + w) of position p, w = w_long, with S = sum z and
// V = w * sum z^2 - S^2 per window (V / w^2: the window's variance in raw units; pivot-free), the t-statistic of
// events.c:338-361 in exact arithmetic is sqrt(w) |S_R - S_L| / sqrt(V_L + V_R), and the reference's rounded value
// t2(p) obeys (oracle/proofs/long_filter.md)
//     t2(p) > thr   ==>   w (|S_R - S_L| + a)^2 (1 + 2^-12)  >=  thr^2 (V_L + V_R) - B,
//     a = 8 u w M,  B = thr^2 64 u w^2 M^2,  u = 2^-24,  M >= max |raw + offset| over both windows.
// The kernel keeps THREE running sums over the 2w samples around p, each updated with the samples that enter,
// cross the middle and leave (exact integers: nothing drifts):
//     dS = S_R - S_L,   T = S_L + S_R,   Q = sum z^2 over both windows,
// and uses  V_L + V_R = w Q - (T^2 + dS^2) / 2:
//     candidate  <=>  NOT( w (1 + 2^-12) (|dS| + a)^2 + thr^2 (T^2 + dS^2) / 2 + B  <  thr^2 w Q ).
// long_candidate() evaluates that; a false result PROVES t2(p) <= thr.
constexpr int ZMAX = 1023;
struct LongK {
    float c0;      // pivot (integer valued)
    float ca;      // a
    float cw;      // w (1 + 2^-12)
    float h2;      // thr^2 / 2
    float t2w;     // thr^2 w
    float cB;      // B
    float thr;     // thr_long (9.0 for DNA and RNA; a context parameter so that tests can make the long detector fire)
};
template <int RNA>
SGW_HD LongK long_consts(int c0, float off, float thr_long) {
    constexpr float w = (float)Cfg<RNA>::w2;
    const float u = 5.9604644775390625e-08f;  // 2^-24
    LongK k;
    k.c0 = (float)c0;
    const float M = fmul(fadd(fabsf(fadd(k.c0, off)), (float)(ZMAX + 2)), 1.0001f);  // >= |raw + offset| while |z| <= ZMAX
    k.ca = fmul(fmul(8.0f * u * w, M), 1.0001f);
    k.cw = w * (1.0f + 0.000244140625f);
    k.thr = thr_long;
    const float thr2 = fmul(thr_long, thr_long);
    k.h2 = fmul(thr2, 0.5f);
    k.t2w = fmul(thr2, w);
    k.cB = fmul(fmul(fmul(thr2, 64.0f * u * w * w), fmul(M, M)), 1.0001f);
    return k;
}
// dS = S_R - S_L, T = S_L + S_R, Q = sum of z^2 over both windows (exact integers)
SGW_HD bool long_candidate(float dS, float T, float Q, const LongK& k) {
    const float g = fadd(fabsf(dS), k.ca);
    const float sq = ffma(T, T, fmul(dS, dS));
    const float lhs = ffma(fmul(g, g), k.cw, ffma(sq, k.h2, k.cB));
    const float rhs = fmul(Q, k.t2w);
    return !(lhs < rhs);  // (NaN / infinity anywhere: candidate)
}

// a sample enters windows whose t-statistics are computed at most two blocks later
constexpr int DIRTY_BLOCKS = 3;  // the block with the sample and the next two hold every window that contains it
static_assert(2 * Cfg<0>::w2 - 1 <= 2 * Cfg<0>::U + 0 && 2 * Cfg<1>::w2 - 1 <= 2 * Cfg<1>::U, "windows reach at most two blocks ahead");

// ---- the register rings --------------------------------------------------------------------------------------------
template <int RNA>
struct Rings {
    using C = Cfg<RNA>;
    double P[C::R1], PQ[C::R1];    // P[j & (R1-1)] = sum of x over the walked samples before j
    float A1[C::R1];               // left-window terms of the short window by window start
    double L1[C::R1];
    float T1c[C::w1];              // t1 of the last w1 positions of the previous block
    // integer side (z = raw - pivot, exact in float): the samples of the previous block, the last ZH2 ones of the
    // block before it, and the three running sums of the long window pair (see long_candidate)
    static constexpr int ZH2 = 2 * C::w2 - C::U;   // how far the 2w-sample history reaches into the block before the previous one
    float Z1[C::U], Z2[ZH2];
    float zdS, zT, zQ;
    SGW_HD void clear() {
#pragma unroll
        for (int k = 0; k < C::R1; k++) { P[k] = 0.0; PQ[k] = 0.0; A1[k] = 0.0f; L1[k] = 0.0; }
#pragma unroll
        for (int k = 0; k < C::w1; k++) T1c[k] = 0.0f;
#pragma unroll
        for (int k = 0; k < C::U; k++) Z1[k] = 0.0f;
#pragma unroll
        for (int k = 0; k < ZH2; k++) Z2[k] = 0.0f;
        zdS = 0.0f; zT = 0.0f; zQ = 0.0f;
    }
};

// One position of the fast path: the short detector (events.c:385-440 for the short one) and the bookkeeping of the
// long detector's lives.
//   cand : long_candidate() of position u (false proves t2(u) <= thr_long)
//   rec  : the step is owned by this chunk (jobs are only created for owned steps)
template <int RNA, bool DIRECT, class E, class Io>
SGW_HD void det_step(WalkDet& d, int m, int u, float c1, bool cand, bool rec, PeakAcc& acc, const E& on_emit, Io& io) {
    bool maskl; int p2;
    const int over = det_one<true, RNA>(d.s_pv, d.s_ps, m, u, c1, thr_short<RNA>(), acc, maskl, p2, on_emit);
    acc.oldest = acc.oldest > over ? acc.oldest : over;
    // the short detector holds a peak above its threshold: the long detector is reset and masked up to
    // peak_pos + w_short (events.c:414-422) -- the current life ends before this step, a new one starts with it
    // (straight-line selects: a rare branch per step measured 14 % slower than these few predicated instructions)
    bool hot = d.l_hot;
    if (maskl & hot & rec) {                                 // rare: the life that ends here may have emitted
        if (DIRECT) io.job(d.l_start, u);
        else { acc.job_ls = d.l_start; acc.job_end = u; acc.jobs++; }   // (parked; a second one redoes the block)
    }
    const int ls = (p2 & ~PS_OPEN) + Cfg<RNA>::w1 + 1;
    d.l_start = maskl ? (ls > u ? ls : u) : d.l_start;
    hot = (hot & !maskl) | ((u >= d.l_start) & cand);        // the long detector is stepped at u (events.c:387)
    d.l_hot = hot;
}

// The rare path of a block: the t-statistics of the whole block from the raw samples with the reference's own
// operations, then the block's detector steps again from the state the block started in; every emitted peak is
// recorded on its own (io.peak), however old it is.
template <int RNA> struct Redo { WalkDet d; PeakAcc acc; float t1c[Cfg<RNA>::w1]; };
template <int RNA, bool EDGE, class Io>
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline
#endif
Redo<RNA> redo_block(Io& io, Redo<RNA> in, int tau0, int n, int sh, bool rec, float off, float unit, float thr_long) {
    using C = Cfg<RNA>;
    constexpr int w1 = C::w1, w2 = C::w2, U = C::U;
    (void)thr_long;
    // the short t-statistic of the block's U positions with the reference's own operations; the window sums slide
    // (three samples per position instead of 2 w1). The long window is not evaluated at all: every position of a
    // redone block simply counts as a candidate (a life that holds one is replayed by long_job, which is exact).
    float e1[U];
    SlidingT<Io> slide;
    for (int m = 0; m < U; m++) e1[m] = slide.next(io, tau0 + m - w1 + 1, w1, n, off, unit);   // (0 where the windows leave the read)
    Redo<RNA> r;
    r.d = in.d; r.acc.mk = 0u; r.acc.oldest = 0; r.acc.jobs = 0;
    const int u0 = tau0 - w2 + 1 + sh;
    for (int m = 0; m < U; m++) {
        const int j2 = tau0 + m - w2 + 1;
        const float c1 = m >= w1 ? e1[m - w1] : in.t1c[m];
        if (!EDGE || (j2 >= 1 && j2 < n)) {
            PeakAcc unused; unused.mk = 0u; unused.oldest = 0; unused.jobs = 0;   // the mask is not used here
            det_step<RNA, true>(r.d, 0, u0 + m, c1, true, rec, unused, [&](int pos) { if (rec) io.peak(pos); }, io);
        }
    }
    for (int k = 0; k < w1; k++) r.t1c[k] = e1[U - w1 + k];
    return r;
}

// One block of U samples, two samples at a time: window terms -> the short t-statistic -> the exact integer sums
// of the long windows and their candidate test -> one step of the short detector, in one straight line of code.
// The same code also fills the rings at the start of a chunk: the two blocks before the first detector step run
// it with live == false (their t-statistics come from partly filled rings and are never used, the detector state
// is reset afterwards), except that the last w1 values of t1 of the second of them ARE the ones the first real
// steps read (live_t1).
// EDGE: the block may touch positions outside the read [0, n): samples there count as 0, t is 0 outside
// w <= i <= n-w (events.c:328-338), the detector only steps positions 1 <= p < n (position 0 is masked: 387).
//   x[m]   : pA of sample tau0 + m (0 outside the read when EDGE)
//   z[m]   : raw - pivot of the same sample (0 outside the read when EDGE)
//   tau0   : read index of the block's first sample, a multiple of U
//   sh     : read_off & 31
//   zdirty : > 0 when a window of this block holds a sample with |z| > ZMAX (its sums may be inexact): every
//            position of the block is a candidate
// dirty > 0: a window of this block holds a LOW sample (pA <= 0 or barely above: a glitch; about 3 reads in 100 of
// real R9.4 data have one). widen_pos does not apply to such a value, so the chunk drivers hand it to this code as
// 0 (it then adds nothing to the running sums), and this block and the next two -- every window that contains the
// sample -- take their t-statistics from redo_block, which works from the raw samples with the reference's own
// operations. Every other window holds unmodified positive samples only, and its sums are differences of running
// sums that miss the same samples on both sides: nothing else changes and the read stays on the fast path.
template <int RNA, bool EDGE, class Io>
SGW_HD void walk_block(Rings<RNA>& g, WalkDet& d, const float (&x)[Cfg<RNA>::U], const float (&z)[Cfg<RNA>::U], int tau0,
                       int n, int sh, bool rec, bool live, bool live_t1, int dirty, int zdirty, float off, float unit,
                       const LongK& lk, Io& io) {
    using C = Cfg<RNA>;
    constexpr int w1 = C::w1, w2 = C::w2, M1 = C::R1 - 1, M2 = C::R2 - 1, U = C::U;
    float t1v[U];
    bool ok = true, ok_t1 = true;  // ok_t1: among the last w1 values of t1
    const WalkDet d0 = d;
    PeakAcc acc;
    acc.mk = 0u; acc.oldest = 0; acc.jobs = 0;
    const int u0 = tau0 - w2 + 1 + sh;                              // position of the block's first step
    float xq[U];                                                    // float squares (events.c:301), two per instruction
#pragma unroll
    for (int m = 0; m < U; m += 2) {
        const F2 p = {x[m], x[m + 1]};
        const F2 q = f2sq(p);
        xq[m] = q.lo; xq[m + 1] = q.hi;
    }
    float zdS = g.zdS, zT = g.zT, zQ = g.zQ;
#pragma unroll
    for (int m0 = 0; m0 < U; m0 += 2) {
        // ---- the short t-statistic of the two positions j1 = tau0 + m - w1 + 1, m = m0, m0 + 1 ----
        float a1[2], b1[2]; double l1[2], v1[2], b1sq[2];
        float al[2]; double ll[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int m = m0 + h;
            // static ring slots: every index below is (m + const) & mask because tau0 is a multiple of U
            const double xd = widen_pos(x[m]);                      // (a sample handed over as 0 widens to 2^-127: it is
            const double qd = widen_pos(xq[m]);                     //  absorbed by the first addition to a real sum)
            const double pn = dadd(g.P[m & M1], EDGE ? (x[m] > 0.0f ? xd : 0.0) : xd);
            const double pqn = dadd(g.PQ[m & M1], EDGE ? (x[m] > 0.0f ? qd : 0.0) : qd);
            g.P[(m + 1) & M1] = pn;
            g.PQ[(m + 1) & M1] = pqn;
            const int s1 = (m - w1 + 1) & M1;                       // slot of j1 (short window [j1, j1+w1))
            const int s1l = (m - 2 * w1 + 1) & M1;                  // slot of j1 - w1
            const double d1 = dsub(pn, g.P[s1]), e1 = dsub(pqn, g.PQ[s1]);
            window_terms<w1>(d1, e1, a1[h], l1[h], b1[h], v1[h], b1sq[h]);
            al[h] = g.A1[s1l]; ll[h] = g.L1[s1l];
        }
        bool k0 = true, k1 = true;
        tstat_two<w1>(al[0], ll[0], b1[0], v1[0], b1sq[0], al[1], ll[1], b1[1], v1[1], b1sq[1], t1v[m0], k0, t1v[m0 + 1], k1);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int m = m0 + h;
            bool kh = h ? k1 : k0;
            const int j1 = tau0 + m - w1 + 1;
            if (EDGE) {  // positions whose windows leave the read do not count: their t is 0
                const bool in1 = (j1 >= w1) & (j1 + w1 <= n);
                kh = kh | !in1;
                t1v[m] = in1 ? t1v[m] : 0.0f;
            }
            if (m >= U - w1) ok_t1 = ok_t1 & kh; else ok = ok & kh;
            const int s1 = (m - w1 + 1) & M1;
            g.A1[s1] = a1[h]; g.L1[s1] = l1[h];
        }
        // ---- the long window's exact integer sums, the candidate test, the detector steps ----
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int m = m0 + h;
            // the newest sample tau0 + m enters the right window, the one w2 back crosses into the left window,
            // the one 2 w2 back leaves (static indices: this block, the previous one, the one before)
            const float zn = z[m];
            const float zm = m >= w2 ? z[m >= w2 ? m - w2 : 0] : g.Z1[m < w2 ? m + U - w2 : 0];
            const float zo = m >= 2 * w2 - U ? g.Z1[m >= 2 * w2 - U ? m + U - 2 * w2 : 0] : g.Z2[m < 2 * w2 - U ? m : 0];
#if defined(WALK_X_NOCAND)   // (timing experiments, tools/run_variants.sh: results are wrong with any WALK_X_* defined)
            zT = fadd(zT, zn); (void)zm; (void)zo;
            bool cand = false;
#else
            zT = fadd(zT, fsub(zn, zo));
            zQ = fadd(zQ, fsub(fmul(zn, zn), fmul(zo, zo)));
            zdS = fadd(fadd(zdS, fadd(zn, zo)), fmul(zm, -2.0f));
            bool cand = long_candidate(zdS, zT, zQ, lk);
#endif
            const int j2 = tau0 + m - w2 + 1;
            if (EDGE) cand = cand & (j2 >= w2) & (j2 + w2 <= n);    // t2 is 0 there (events.c:328-338)
            cand = cand | (zdirty > 0);
            const float c1 = m >= w1 ? t1v[m >= w1 ? m - w1 : 0] : g.T1c[m < w1 ? m : 0];  // t1(j2), computed w1 samples ago
#if defined(WALK_X_NODET)
            d.s_pv = fadd(d.s_pv, c1); d.l_hot = d.l_hot | cand;
#elif defined(WALK_X_NOLIFE)
            { bool mk_; int p2_; const int over_ = det_one<true, RNA>(d.s_pv, d.s_ps, m, u0 + m, c1, thr_short<RNA>(), acc, mk_, p2_, NoEmit());
              acc.oldest = acc.oldest > over_ ? acc.oldest : over_; d.l_hot = d.l_hot | (cand & mk_); d.l_start += p2_ & 1; }
#else
            if (!EDGE || (j2 >= 1 && j2 < n)) det_step<RNA, false>(d, m, u0 + m, c1, cand, rec, acc, NoEmit(), io);
#endif
        }
    }
    g.zdS = zdS; g.zT = zT; g.zQ = zQ;
#pragma unroll
    for (int k = 0; k < Rings<RNA>::ZH2; k++) g.Z2[k] = g.Z1[k + U - Rings<RNA>::ZH2];
#pragma unroll
    for (int k = 0; k < U; k++) g.Z1[k] = z[k];
#if defined(WALK_TEST_REDO)  // host tests: take the rare path on every third block
    ok = ok & ((tau0 / U) % 3 != 0);
#endif
    if (dirty > 0) { ok = false; ok_t1 = false; }  // a window of this block holds a LOW sample
    // rare (about 3 blocks in 100,000): a t-statistic next to a rounding midpoint, or a peak too old for the mask
    if ((!ok & live) | (!ok_t1 & live_t1) | (rec & ((acc.oldest > PkCfg<RNA>::FAR) | (acc.jobs > 1)))) {
        Redo<RNA> in;
        in.d = d0; in.acc = acc;
#pragma unroll
        for (int k = 0; k < w1; k++) in.t1c[k] = g.T1c[k];
        const Redo<RNA> r = redo_block<RNA, EDGE>(io, in, tau0, n, sh, rec, off, unit, lk.thr);
        d = r.d; acc = r.acc;
#pragma unroll
        for (int k = 0; k < w1; k++) t1v[U - w1 + k] = r.t1c[k];
    }
    if (rec && acc.mk) io.peaks32(u0 - PkCfg<RNA>::LEAD, acc.mk);  // peaks are owned by the step that emits them
    if (acc.jobs == 1) io.job(acc.job_ls, acc.job_end);            // (0 after redo_block: it creates its jobs itself)
#pragma unroll
    for (int k = 0; k < w1; k++) g.T1c[k] = t1v[U - w1 + k];
}

// ---- chunk drivers -------------------------------------------------------------------------------------------------
// Every read is cut into chunks of L samples; the last chunk takes the remainder (1..L samples), a read of at
// most L samples is a single chunk. `Io` supplies the memory side (device: walk.cu, host checker:
// tests/tools/host_walk.cpp):
//   load8(t, v)        the four 32-bit words holding samples [t, t+8) of the read (t multiple of 8)
//   want_pa() / store_pa8(t, x) / store_pa1(t, x)
//   peaks32(ub, mk)    record the emitted peaks of a block: bit k of mk (nonzero) <=> a peak at shifted position ub + k
//   peak(pos)          record one peak
//   put_begin(c) / put_end(c)   the chunk's canonical detector state after the warm-up / after its last step
//   job(l_start, end)  a life of the long detector that may emit: its steps [l_start, end) (shifted positions;
//                      l_start == LS_PRED: the life began in an earlier chunk, end == LS_CONT: it runs past this one)
//   witness(rmin, rmax, low_t)  extreme raw values of samples of the read (any superset of the owned samples);
//                               samples with raw <= low_t are LOW (reported through low_samples by their blocks)
//   low_samples(t, lo, hi)      LOW samples in the group of 8 that holds read index t: their smallest nonzero /
//                               largest |pA| bit patterns (lo > hi: all of them are zero)
SGW_HD uint32_t n_chunks(uint32_t n, uint32_t L) { return (n + L - 1u) / L; }

// int16 -> float without a conversion instruction: with the sign bit flipped, the 16 bits are raw + 32768 in
// [0, 65535]; placed under the exponent of 2^23 they read as the float 2^23 + raw + 32768, and subtracting
// 2^23 + 32768 is exact. Then misc.c:28: float add of the offset, float multiply by the unit (two samples per
// instruction); z = raw - pivot from the same exact float (one more packed subtraction).
SGW_HD void cvt8(const int (&v)[4], float off, float unit, float c0, float* x, float* z) {
    const F2 bias = {8421376.0f, 8421376.0f}, off2 = {off, off}, unit2 = {unit, unit}, c2 = {c0, c0};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t w = (uint32_t)v[k] ^ 0x80008000u;
        F2 r;
        r.lo = bits_f(0x4b000000u | (w & 0xffffu));
        r.hi = bits_f(0x4b000000u | (w >> 16));
        r = f2sub(r, bias);
        const F2 zz = f2sub(r, c2);
        r = f2mul(f2add(r, off2), unit2);
        x[2 * k] = r.lo;
        x[2 * k + 1] = r.hi;
        z[2 * k] = zz.lo;
        z[2 * k + 1] = zz.hi;
    }
}
SGW_HD void cvt8(const int (&v)[4], float off, float unit, float* x) {  // pA only (sequential-order kernels, pa_kernel)
    const F2 bias = {8421376.0f, 8421376.0f}, off2 = {off, off}, unit2 = {unit, unit};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t w = (uint32_t)v[k] ^ 0x80008000u;
        F2 r;
        r.lo = bits_f(0x4b000000u | (w & 0xffffu));
        r.hi = bits_f(0x4b000000u | (w >> 16));
        r = f2mul(f2add(f2sub(r, bias), off2), unit2);
        x[2 * k] = r.lo;
        x[2 * k + 1] = r.hi;
    }
}
#if defined(__CUDA_ARCH__)
SGW_HD uint32_t min_s16x2(uint32_t a, uint32_t b) { return __vmins2(a, b); }
SGW_HD uint32_t max_s16x2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
#else
SGW_HD uint32_t min_s16x2(uint32_t a, uint32_t b) {
    const int16_t al = (int16_t)(a & 0xffff), ah = (int16_t)(a >> 16), bl = (int16_t)(b & 0xffff), bh = (int16_t)(b >> 16);
    return (uint32_t)(uint16_t)(al < bl ? al : bl) | ((uint32_t)(uint16_t)(ah < bh ? ah : bh) << 16);
}
SGW_HD uint32_t max_s16x2(uint32_t a, uint32_t b) {
    const int16_t al = (int16_t)(a & 0xffff), ah = (int16_t)(a >> 16), bl = (int16_t)(b & 0xffff), bh = (int16_t)(b >> 16);
    return (uint32_t)(uint16_t)(al > bl ? al : bl) | ((uint32_t)(uint16_t)(ah > bh ? ah : bh) << 16);
}
#endif

#if defined(__CUDA_ARCH__)
SGW_HD bool any_le_s16x2(uint32_t a, uint32_t b) { return __vcmples2(a, b) != 0u; }  // a.lo <= b.lo || a.hi <= b.hi
#else
SGW_HD bool any_le_s16x2(uint32_t a, uint32_t b) {
    return (int16_t)(a & 0xffff) <= (int16_t)(b & 0xffff) || (int16_t)(a >> 16) <= (int16_t)(b >> 16);
}
#endif
// LOW samples: raw <= t, where t is LOW_MARGIN above the largest raw value whose pA is not positive
// (fl((float)raw + off) <= 0; the unit is positive). They include every sample with pA <= 0 -- glitches -- and the
// few levels just above zero, which no pore produces (t + 1 maps to about 11 pA for R9.4 scaling). The chunk drivers
// zero them for the running sums (see walk_block), mark the blocks around them dirty and report their magnitudes
// on their own, so that the exact-sum witness of all OTHER samples can start at pA(t + 1) instead of at the
// smallest nonzero magnitude a range could hold.
// Returns t clamped to int16; *can is false when no int16 value is low.
constexpr int LOW_MARGIN = 64;
SGW_HD int low_threshold(float off, bool* can) {
    float tf = floorf(-off);
    tf = tf < -40000.0f ? -40000.0f : tf > 40000.0f ? 40000.0f : tf;
    int t = (int)tf;
    for (int k = 0; k < 4 && fadd((float)(t + 1), off) <= 0.0f; k++) t++;
    for (int k = 0; k < 4 && fadd((float)t, off) > 0.0f; k++) t--;
    t += LOW_MARGIN;
    *can = t >= -32768;
    return t > 32767 ? 32767 : t < -32768 ? -32768 : t;
}
SGW_HD uint32_t pack_s16x2(int t) { return ((uint32_t)t & 0xffffu) * 0x10001u; }
SGW_HD int s16_lo(uint32_t v) { return (int)(int16_t)(v & 0xffffu); }
SGW_HD int s16_hi(uint32_t v) { return (int)v >> 16; }

// the pivot of a chunk: the median of the first three samples it walks (one glitch sample cannot drag it away)
SGW_HD int pivot_of(int a, int b, int c) {
    const int lo = a < b ? a : b, hi = a < b ? b : a;
    return c < lo ? lo : c > hi ? hi : c;
}

// the LOW samples of one group of 8: their magnitudes go to the witness, their values become 0 for the running sums
template <class Io>
SGW_HD void zero_low8(Io& io, int t, const int (&v)[4], int low_t, float* x) {
    uint32_t lo = 0xffffffffu, hi = 0u;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const int raw = (q & 1) ? (v[q >> 1] >> 16) : (int)(int16_t)(v[q >> 1] & 0xffff);
        if (raw <= low_t) {
            const uint32_t a = f_bits(x[q]) & 0x7fffffffu;
            if (a != 0u) { lo = a < lo ? a : lo; hi = a > hi ? a : hi; }
            x[q] = 0.0f;
        }
    }
    io.low_samples(t, lo, hi);
}

// what a chunk driver needs to turn the packed samples of one block into x[] / z[] on the unchecked path
struct BlockCvt {
    bool can_low, zlo_on, zhi_on, pa;
    int low_t;
    uint32_t low2, zlo2, zhi2;
};
SGW_HD BlockCvt block_cvt(float off, int c0, bool want_pa) {
    BlockCvt k;
    k.low_t = low_threshold(off, &k.can_low);
    k.low2 = pack_s16x2(k.low_t);
    // |z| <= ZMAX  <=>  c0 - ZMAX <= raw <= c0 + ZMAX (clamped so that the packed comparisons cannot wrap)
    k.zlo2 = pack_s16x2(c0 - ZMAX - 1 < -32768 ? -32768 : c0 - ZMAX - 1);
    k.zhi2 = pack_s16x2(c0 + ZMAX + 1 > 32767 ? 32767 : c0 + ZMAX + 1);
    k.zlo_on = c0 - ZMAX - 1 >= -32768; k.zhi_on = c0 + ZMAX + 1 <= 32767;
    k.pa = want_pa;
    return k;
}
// one group of 8 samples (all inside the read): pA, z, the optional pA store, LOW samples, packed extremes
template <class Io>
SGW_HD void cvt_group(Io& io, const BlockCvt& k, const LongK& lk, const int (&v)[4], int t8, bool own, float off, float unit,
                      float* x, float* z, uint32_t& bmin, uint32_t& bmax, int& dirty) {
    const uint32_t m8 = min_s16x2(min_s16x2((uint32_t)v[0], (uint32_t)v[1]), min_s16x2((uint32_t)v[2], (uint32_t)v[3]));
    bmin = min_s16x2(bmin, m8);
    bmax = max_s16x2(max_s16x2(bmax, (uint32_t)v[0]), max_s16x2((uint32_t)v[1], max_s16x2((uint32_t)v[2], (uint32_t)v[3])));
    cvt8(v, off, unit, lk.c0, x, z);
    if (k.pa && own) io.store_pa8(t8, x);
    if (k.can_low & any_le_s16x2(m8, k.low2)) {   // LOW samples in this group (rare)
        dirty = DIRTY_BLOCKS;
        zero_low8(io, t8, v, k.low_t, x);
    }
}
// after the groups of a block: samples far from the pivot (rare) enter the integer sums clamped, and every window
// that holds one is a candidate of the long detector's test
template <int U>
SGW_HD void clamp_far(const BlockCvt& k, uint32_t bmin, uint32_t bmax, float* z, int& zdirty) {
    if ((k.zlo_on & any_le_s16x2(bmin, k.zlo2)) | (k.zhi_on & any_le_s16x2(k.zhi2, bmax))) {  // |z| > ZMAX
        zdirty = DIRTY_BLOCKS;
#pragma unroll
        for (int q = 0; q < U; q++)
            z[q] = z[q] > (float)(ZMAX + 1) ? (float)(ZMAX + 1) : z[q] < -(float)(ZMAX + 1) ? -(float)(ZMAX + 1) : z[q];
    }
}

// how many blocks ahead of its loads a chunk asks for its samples' cache line (io.prefetch; measured on the bench
// batch, profiles/r02_variants_pf.txt: none 3.77 ms, 2-4 blocks 3.71, 8-16 blocks 3.62, 32 blocks 3.66)
#if !defined(WALK_PF_BLOCKS)
#define WALK_PF_BLOCKS 12
#endif

// interior chunk k (1 <= k <= nch-2) of a read: every access is inside the read, no bounds checks.
// ONE loop (one copy of the block code in the instruction cache) runs the two ring-fill blocks, the detector
// warm-up and the owned samples.
template <int RNA, class Io>
SGW_HD void walk_interior(Io& io, int n, float off, float unit, int sh, int L, int W, int k, float thr_long) {
    using C = Cfg<RNA>;
    constexpr int U = C::U;
    const int s0 = k * L, s1 = s0 + L;               // the samples this chunk owns
    const int t_live = s0 - W;                       // block whose first step is the first real detector step
    Rings<RNA> g;
    g.clear();
    WalkDet d;
    det_cold(d, 0);
    float x[U], z[U];
    int v[4];
    uint32_t vmin = 0x7fff7fffu, vmax = 0x80008000u;  // packed int16 min / max (over warm-up and owned samples)
    int dirty = 0, zdirty = 0;
    int vn[U / 8][4];                                 // the next block's samples, loaded one block ahead
#pragma unroll
    for (int h = 0; h < U / 8; h++) io.load8(t_live - C::FILL * U + 8 * h, vn[h]);
    const int c0 = pivot_of(s16_lo((uint32_t)vn[0][0]), s16_hi((uint32_t)vn[0][0]), s16_lo((uint32_t)vn[0][1]));
    const LongK lk = long_consts<RNA>(c0, off, thr_long);
    const BlockCvt bc = block_cvt(off, c0, io.want_pa());
#pragma unroll 1
    for (int tau = t_live - C::FILL * U; tau < s1; tau += U) {
        const bool own = tau >= s0;
        if (tau == t_live) det_cold(d, t_live - C::LAG + sh);   // forget the steps taken on partly filled rings
        if (tau == s0) {
            io.put_begin(canon_of<RNA>(d, s0 - C::LAG + sh));
            d.l_start = LS_PRED; d.l_hot = false;               // the running life began before this chunk's steps
        }
        const int tn = tau + U < s1 ? tau + U : tau;  // (the last block is simply loaded again)
        io.prefetch(tau + WALK_PF_BLOCKS * U < s1 ? tau + WALK_PF_BLOCKS * U : tau);
        uint32_t bmin = 0x7fff7fffu, bmax = 0x80008000u;  // packed extremes of this block's samples
#pragma unroll
        for (int h = 0; h < U / 8; h++) {
#pragma unroll
            for (int q = 0; q < 4; q++) v[q] = vn[h][q];
            io.load8(tn + 8 * h, vn[h]);
            cvt_group(io, bc, lk, v, tau + 8 * h, own, off, unit, x + 8 * h, z + 8 * h, bmin, bmax, dirty);
        }
        vmin = min_s16x2(vmin, bmin);
        vmax = max_s16x2(vmax, bmax);
        clamp_far<U>(bc, bmin, bmax, z, zdirty);
        walk_block<RNA, false>(g, d, x, z, tau, n, sh, own, tau >= t_live, tau >= t_live - U, dirty, zdirty, off, unit, lk, io);
        dirty = dirty > 0 ? dirty - 1 : 0;
        zdirty = zdirty > 0 ? zdirty - 1 : 0;
    }
    if (d.l_hot) io.job(d.l_start, LS_CONT);
    io.put_end(canon_of<RNA>(d, s1 - C::LAG + sh));
    const int rmin0 = s16_lo(vmin), rmin1 = s16_hi(vmin);
    const int rmax0 = s16_lo(vmax), rmax1 = s16_hi(vmax);
    io.witness(rmin0 < rmin1 ? rmin0 : rmin1, rmax0 > rmax1 ? rmax0 : rmax1, bc.can_low ? bc.low_t : -32769);
}

// first chunk (last == 0) or last chunk (last == 1; only when the read has >= 2 chunks) of a read.
// Only the blocks at either end of the read need bounds checks (samples or windows outside the read); the blocks in
// between are the interior chunks' unchecked code. The chunk is walked in three phases -- checked head, unchecked
// middle, checked tail -- each a function of its own (not inlined on the device: one loop holding both block variants
// spilled ~100 registers and ran its blocks twice as slowly as an interior chunk; the phases pass the walk's state
// through EdgeCtx, three calls per chunk).
#if defined(__CUDACC__)
#define SGW_PHASE __host__ __device__ __noinline__
#else
#define SGW_PHASE inline
#endif
template <int RNA>
struct EdgeCtx {
    Rings<RNA> g;
    WalkDet d;
    int rmin, rmax, dirty, zdirty;
    int n, sh, s0, s1, t_live, last, low_t, c0;
    float off, unit;
    bool pa;
    LongK lk;
    BlockCvt bc;
};
// what every block of an edge chunk does before its samples
template <int RNA, class Io>
SGW_HD void edge_block_begin(Io& io, EdgeCtx<RNA>& e, int tau) {
    using C = Cfg<RNA>;
    if (e.last && tau == e.t_live) det_cold(e.d, e.t_live - C::LAG + e.sh);
    if (e.last && tau == e.s0) {
        io.put_begin(canon_of<RNA>(e.d, e.s0 - C::LAG + e.sh));
        e.d.l_start = LS_PRED; e.d.l_hot = false;
    }
}
// blocks [tau_from, tau_to) whose samples and windows all lie inside the read: the interior chunks' code
template <int RNA, class Io>
SGW_PHASE void edge_inner(Io& io, EdgeCtx<RNA>& ec, int tau_from, int tau_to) {
    using C = Cfg<RNA>;
    constexpr int U = C::U;
    // every lane of the warp takes as many turns as the lane with the most blocks and sits out the ones it has no
    // block for: the lanes leave the loop, and the function, together
    const int turns = io.warp_max(tau_to > tau_from ? (tau_to - tau_from) / U : 0);
    EdgeCtx<RNA> e = ec;                              // (registers for the duration of the phase)
    float x[U], z[U];
    int vn[U / 8][4];                                 // the next block's samples, loaded one block ahead
    if (tau_from < tau_to) {
#pragma unroll
        for (int h = 0; h < U / 8; h++) io.load8(tau_from + 8 * h, vn[h]);
    }
    int tau = tau_from;
#pragma unroll 1
    for (int turn = 0; turn < turns; turn++, tau += U) {
        if (tau >= tau_to) continue;
        const bool own = tau >= e.s0;
        edge_block_begin<RNA>(io, e, tau);
        const int tn = tau + U < tau_to ? tau + U : tau;   // (the last block is simply loaded again)
        io.prefetch(tau + WALK_PF_BLOCKS * U < tau_to ? tau + WALK_PF_BLOCKS * U : tau);
        uint32_t bmin = 0x7fff7fffu, bmax = 0x80008000u;
#pragma unroll
        for (int h = 0; h < U / 8; h++) {
            int v[4];
#pragma unroll
            for (int q = 0; q < 4; q++) v[q] = vn[h][q];
            io.load8(tn + 8 * h, vn[h]);
            cvt_group(io, e.bc, e.lk, v, tau + 8 * h, own, e.off, e.unit, x + 8 * h, z + 8 * h, bmin, bmax, e.dirty);
        }
        if (own) {  // (an owned inner block lies inside [s0, s1): s1 is n or a multiple of U)
            const int b0 = s16_lo(bmin), b1 = s16_hi(bmin), c0x = s16_lo(bmax), c1x = s16_hi(bmax);
            e.rmin = b0 < e.rmin ? b0 : e.rmin; e.rmin = b1 < e.rmin ? b1 : e.rmin;
            e.rmax = c0x > e.rmax ? c0x : e.rmax; e.rmax = c1x > e.rmax ? c1x : e.rmax;
        }
        clamp_far<U>(e.bc, bmin, bmax, z, e.zdirty);
        walk_block<RNA, false>(e.g, e.d, x, z, tau, e.n, e.sh, own, tau >= e.t_live, tau >= e.t_live - U, e.dirty, e.zdirty,
                               e.off, e.unit, e.lk, io);
        e.dirty = e.dirty > 0 ? e.dirty - 1 : 0;
        e.zdirty = e.zdirty > 0 ? e.zdirty - 1 : 0;
    }
    ec = e;
}
// blocks [tau_from, tau_to) at either end of the read: every sample and window is checked against the read's bounds
template <int RNA, class Io>
SGW_PHASE void edge_checked(Io& io, EdgeCtx<RNA>& ec, int tau_from, int tau_to) {
    using C = Cfg<RNA>;
    constexpr int U = C::U;
    const int turns = io.warp_max(tau_to > tau_from ? (tau_to - tau_from) / U : 0);   // (see edge_inner)
    EdgeCtx<RNA> e = ec;
    float x[U], z[U];
    const int n = e.n, c0 = e.c0, low_t = e.low_t;
    int tau = tau_from;
#pragma unroll 1
    for (int turn = 0; turn < turns; turn++, tau += U) {
        if (tau >= tau_to) continue;
        const bool own = tau >= e.s0;
        edge_block_begin<RNA>(io, e, tau);
#pragma unroll
        for (int h = 0; h < U / 8; h++) {
            const int t8 = tau + 8 * h;
            int v[4] = {0, 0, 0, 0};
            if (t8 < n) io.load8(t8, v);  // t8 < n: inside the read's padded span
            float y[8], zy[8];
            cvt8(v, e.off, e.unit, e.lk.c0, y, zy);
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const bool in = t8 + q < n;
                const int raw = (q & 1) ? (v[q >> 1] >> 16) : (int)(int16_t)(v[q >> 1] & 0xffff);
                const bool low = in & (raw <= low_t);   // rare: handed to the block as 0, its windows are dirty
                x[8 * h + q] = (in & !low) ? y[q] : 0.0f;
                const bool zbad = in & ((raw - c0 > ZMAX) | (c0 - raw > ZMAX));  // rare: enters the integer sums clamped
                z[8 * h + q] = !in ? 0.0f : !zbad ? zy[q] : raw > c0 ? (float)(ZMAX + 1) : -(float)(ZMAX + 1);
                if (zbad) e.zdirty = DIRTY_BLOCKS;
                if (low) {
                    e.dirty = DIRTY_BLOCKS;
                    const uint32_t a = f_bits(y[q]) & 0x7fffffffu;
                    io.low_samples(t8 + q, a != 0u ? a : 0xffffffffu, a);
                }
                if (own && in && t8 + q < e.s1) {
                    e.rmin = raw < e.rmin ? raw : e.rmin;
                    e.rmax = raw > e.rmax ? raw : e.rmax;
                    if (e.pa) io.store_pa1(t8 + q, y[q]);
                }
            }
        }
        walk_block<RNA, true>(e.g, e.d, x, z, tau, n, e.sh, own, tau >= e.t_live, tau >= e.t_live - U, e.dirty, e.zdirty,
                              e.off, e.unit, e.lk, io);
        e.dirty = e.dirty > 0 ? e.dirty - 1 : 0;
        e.zdirty = e.zdirty > 0 ? e.zdirty - 1 : 0;
    }
    ec = e;
}

// The lanes of a warp walk the first (or the last) chunks of 32 different reads and spend different numbers of blocks
// in the middle phase. Every phase therefore runs io.warp_max(blocks) turns in every lane (a lane without a block sits
// the turn out), so that the lanes leave each phase together: lanes that returned from the middle phase one after the
// other ran their checked tails on their own, two lanes active -- 8 times the instructions of the converged tail.
// For that, no lane leaves walk_edge() early: a read without this chunk walks empty phases.
template <int RNA, class Io>
SGW_HD void walk_edge(Io& io, int n, float off, float unit, int sh, int L, int W, int last, float thr_long) {
    using C = Cfg<RNA>;
    constexpr int U = C::U;
    const uint32_t nch = n_chunks((uint32_t)n, (uint32_t)L);
    const bool none = nch == 0u || (last && nch < 2u);   // the read has no such chunk
    // owned samples [s0, s1); owned detector steps [s0 - LAG, s1 - LAG), all remaining steps for the read's last chunk
    EdgeCtx<RNA> e;
    e.n = n; e.sh = sh; e.off = off; e.unit = unit; e.last = last;
    e.s0 = last ? (int)(nch - 1u) * L : 0;
    const bool to_end = last || nch == 1u;
    e.s1 = to_end ? n : L;
    e.t_live = last ? e.s0 - W : 0;                  // >= 2U for a last chunk because L >= W + 2U
    const int step_end = to_end ? n : e.s1 - C::LAG; // owned blocks: until every owned step has been taken
    e.g.clear();
    det_cold(e.d, sh + 1);  // first chunk: the reference's initial state, masked_to = 0: position 0 is skipped (events.c:516-536, 387)
    e.rmin = 32767; e.rmax = -32768;
    bool can_low;
    const int low_t0 = low_threshold(off, &can_low);
    e.low_t = can_low ? low_t0 : -32769;
    e.dirty = 0; e.zdirty = 0;
    e.pa = io.want_pa();
    const int t_first = last ? e.t_live - C::FILL * U : 0;
    e.c0 = 0;
    if (!none) {
        int v0[4] = {0, 0, 0, 0};
        io.load8(t_first, v0);                        // t_first < n: inside the read's padded span
        const int a = s16_lo((uint32_t)v0[0]), b = t_first + 1 < n ? s16_hi((uint32_t)v0[0]) : a;
        e.c0 = pivot_of(a, b, t_first + 2 < n ? s16_lo((uint32_t)v0[1]) : a);
    }
    e.lk = long_consts<RNA>(e.c0, off, thr_long);
    e.bc = block_cvt(off, e.c0, e.pa);
    // blocks tau = t_first, t_first + U, ... while tau - LAG < step_end (all multiples of U); a block is an inner one
    // when INNER_MIN <= tau and tau + U <= n: checked head, unchecked middle, checked tail
    constexpr int INNER_MIN = (2 * C::w2 - 1 + U - 1) / U * U;
    const int tau_end = none ? t_first : (step_end + C::LAG + U - 1) / U * U;
    int in0 = t_first > INNER_MIN ? t_first : INNER_MIN;
    in0 = in0 < tau_end ? in0 : tau_end;
    int in1 = n / U * U;
    in1 = in1 < tau_end ? in1 : tau_end;
    in1 = in1 > in0 ? in1 : in0;
    edge_checked<RNA>(io, e, t_first, in0);
    edge_inner<RNA>(io, e, in0, in1);
    edge_checked<RNA>(io, e, in1, tau_end);
    if (none) return;
    if (to_end) {
        if (e.d.l_hot) io.job(e.d.l_start, n + sh);  // the life ends with the read
    } else {
        if (e.d.l_hot) io.job(e.d.l_start, LS_CONT);
        io.put_end(canon_of<RNA>(e.d, e.s1 - C::LAG + sh));
    }
    if (!last) io.peak(sh);  // event 0 starts at the read's first sample (events.c:490-497)
    if (e.rmin <= e.rmax) io.witness(e.rmin, e.rmax, e.low_t);
}

// ---- a life of the long detector, replayed with the reference's own operations --------------------------------------
// Steps [l_start, end) of the long detector from its reset state (shifted positions; chunk k of the read created
// the job). `Jo` supplies, besides load8 / peak:
//   end_lstart(kk)     word 6 of the END record of chunk kk (its l_start after its last owned step)
//   end_short(kk, pv, ps)   the short detector's state in the same record
// l_start == LS_PRED: the life began in an earlier chunk -- the END records of the chunks before k are searched
// backwards for the last reset (the first chunk always knows: the reference's initial state). end == LS_CONT: the
// life runs past chunk k's last owned step -- from there on the short detector is replayed as well (from chunk
// k's END record, with exact t1) until it resets the long one, which ends the life.
template <int RNA, class Jo>
SGW_HD void long_job(Jo& io, int n, int sh, float off, float unit, int L, int k, int l_start, int end, float thr_long) {
    using C = Cfg<RNA>;
    constexpr int w1 = C::w1, w2 = C::w2;
    for (int kk = k - 1; l_start == LS_PRED && kk >= 0; kk--) l_start = io.end_lstart(kk);
    if (l_start == LS_PRED) return;  // (cannot happen: chunk 0 starts from a concrete state)
    const int stop = n + sh;         // steps exist for positions 1 .. n-1
    const int own_end = end == LS_CONT ? (k + 1) * L - C::LAG + sh : (end < stop ? end : stop);
    float pv = FLT_MAX; int ps = PS_NONE;
    PeakAcc unused; unused.mk = 0u; unused.oldest = 0;
    bool b2; int p2;
    // (one thread per life: the statistics of four consecutive steps are formed before the detector takes them, so
    //  that their dependent chains overlap -- the longest life of a batch is the kernel's duration)
    SlidingT<Jo> slide;
    int u = l_start;
    for (; u + 4 <= own_end; u += 4) {
        float t[4];
#pragma unroll
        for (int q = 0; q < 4; q++) t[q] = slide.next(io, u + q - sh, w2, n, off, unit);
#pragma unroll
        for (int q = 0; q < 4; q++)
            det_one<false, RNA>(pv, ps, 0, u + q, t[q], thr_long, unused, b2, p2, [&](int pos) { io.peak(pos); });
    }
    for (; u < own_end; u++)
        det_one<false, RNA>(pv, ps, 0, u, slide.next(io, u - sh, w2, n, off, unit), thr_long, unused, b2, p2,
                            [&](int pos) { io.peak(pos); });
    if (end != LS_CONT) return;
    float spv; int sps;
    io.end_short(k, &spv, &sps);
    u = own_end;                     // (l_start <= own_end: the life was alive at the chunk's last owned step)
    SlidingT<Jo> slide1;             // the short window from here on, the long one goes on sliding
    for (; u < stop; u++) {
        const float t1 = slide1.next(io, u - sh, w1, n, off, unit);
        const float t2 = slide.next(io, u - sh, w2, n, off, unit);
        det_one<true, RNA>(spv, sps, 0, u, t1, thr_short<RNA>(), unused, b2, p2, NoEmit());
        if (b2) return;              // reset: the life ended before the long detector's step at u
        det_one<false, RNA>(pv, ps, 0, u, t2, thr_long, unused, b2, p2, [&](int pos) { io.peak(pos); });
    }
}

}  // namespace walk
}  // namespace sgpu
