// common.cuh -- shared device helpers for the sigtk B200 hot path (sm_100a).
//
// Every arithmetic step that the reference performs is written with an explicit
// round-to-nearest intrinsic so that nvcc can neither contract (no FMA: the
// reference is built -std=c99 => -ffp-contract=off, Makefile:5) nor reassociate.
// The translation unit is additionally compiled with --fmad=false.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

namespace sgpu {

// detector parameters, reference src/events.c:35-54
struct DetParams {
    int w1, w2;        // short / long window length
    float thr1, thr2;  // thresholds
    float height;      // peak_height
};

__host__ __device__ inline DetParams det_params(int rna) {
    DetParams p;
    if (rna) { p.w1 = 7; p.w2 = 14; p.thr1 = 2.5f; p.thr2 = 9.0f; p.height = 1.0f; }
    else     { p.w1 = 3; p.w2 = 6;  p.thr1 = 1.4f; p.thr2 = 9.0f; p.height = 0.2f; }
    return p;
}

// One batch of reads resident in HBM (see include/sigtk_b200.h for the layout).
struct DevBatch {
    const int16_t*  samples;
    const uint64_t* read_off;  // [n_reads+1]
    const uint32_t* read_len;  // [n_reads]
    const float*    offset;    // [n_reads] (float)rec->offset
    const float*    unit;      // [n_reads] (float)range/(float)digitisation
    uint32_t n_reads;
    int      rna;
    uint64_t span;             // read_off[n_reads]
};

// pA of one sample, reference src/misc.c:28: float add, then float multiply.
__device__ __forceinline__ float pa_of(int16_t raw, float off, float unit) {
    return __fmul_rn(__fadd_rn((float)raw, off), unit);
}

// The windowed t-statistic exactly as the reference evaluates it
// (src/events.c:338-361) given the four window sums in double:
//   sum1/ssq1 : left window  [i-w, i)   (kept in double)
//   sum2/ssq2 : right window [i, i+w)   (narrowed to float at once)
__device__ __forceinline__ float tstat_reference_chain(double sum1, double ssq1, double sum2d,
                                                       double ssq2d, float wf) {
    const double wd = (double)wf;
    const float sum2 = __double2float_rn(sum2d);
    const float ssq2 = __double2float_rn(ssq2d);
    const float mean1 = __double2float_rn(__ddiv_rn(sum1, wd));
    const float mean2 = __fdiv_rn(sum2, wf);
    const float m1sq = __fmul_rn(mean1, mean1);
    const float m2sq = __fmul_rn(mean2, mean2);
    const float v2 = __fdiv_rn(ssq2, wf);
    double acc = __ddiv_rn(ssq1, wd);
    acc = __dsub_rn(acc, (double)m1sq);
    acc = __dadd_rn(acc, (double)v2);
    acc = __dsub_rn(acc, (double)m2sq);
    const float cv = fmaxf(__double2float_rn(acc), FLT_MIN);
    const float delta = __fsub_rn(mean2, mean1);
    const float scaled = __fdiv_rn(cv, wf);
    return __double2float_rn(__ddiv_rn(fabs((double)delta), __dsqrt_rn((double)scaled)));
}

// ---- exact shortcuts (validated on the CPU by oracle/proofs/*.c) ------------------------------------------------
// a / W for the window lengths W in {3,6,7,14}: reciprocal multiply + one FMA residual correction (Markstein).
// Equal to the IEEE quotient for every double the path can produce (sums of floats: never denormal as doubles,
// never -0), and for every float with |a| >= 2^-125 or a == +0 (exhaustive check); smaller floats take the
// IEEE division.
template <int W>
__device__ __forceinline__ double div_w(double a) {
    constexpr double r = 1.0 / (double)W;
    const double q0 = __dmul_rn(a, r);
    const double e = __fma_rn(-(double)W, q0, a);
    return __fma_rn(e, r, q0);
}
template <int W, bool GUARD = true>
__device__ __forceinline__ float div_w(float a) {
    constexpr float r = 1.0f / (float)W;
    const float q0 = __fmul_rn(a, r);
    const float e = __fmaf_rn(-(float)W, q0, a);
    float q = __fmaf_rn(e, r, q0);
    if (GUARD) {  // callers drop the guard only when they know |a| >= 2^-125 or a == +0
        const uint32_t b = __float_as_uint(a);
        if ((b & 0x7fffffffu) < 0x02000000u && b != 0u) q = __fdiv_rn(a, (float)W);
    }
    return q;
}

// (float)(fabs((double)delta) / sqrt((double)scaled))  (events.c:360) through a 22-bit reciprocal square root
// and one third-order correction in double. The product is within a few ulp(double) of the true quotient, so it
// rounds to the same float as the reference's doubly rounded value unless it lies next to a float rounding
// midpoint (or outside the normal float range); those values (about 2 in a million) take the IEEE sqrt + division.
__device__ __forceinline__ float tstat_tail(float delta, float scaled) {
    const double c = (double)scaled;
    float y0f;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0f) : "f"(scaled));  // 2 ulp; denormal inputs are caught below
    const double y0 = (double)y0f;
    const double t = __dmul_rn(c, y0);
    const double e = __fma_rn(-t, y0, 1.0);
    const double p = __fma_rn(0.375, e, 0.5);
    const double ye = __dmul_rn(y0, e);
    const double y = __fma_rn(ye, p, y0);
    const double q = __dmul_rn(fabs((double)delta), y);
    const uint32_t lo = (uint32_t)__double2loint(q), hi = (uint32_t)__double2hiint(q);
    // accept when 2^-126 <= q < 2^126 and the 29 bits below float precision are not within 512 of the midpoint
    const bool in_range = ((hi & 0x7ff00000u) - ((1023u - 126u) << 20)) < (252u << 20);
    const bool off_mid = (((lo & 0x1fffffffu) - (0x10000000u - 512u))) >= 1024u;
    // delta == 0 gives exactly +0 on both routes (scaled > 0); it is common in quantised data, so accept it here
    if (!((in_range && off_mid && scaled >= 1.0e-30f) || delta == 0.0f))
        return __double2float_rn(__ddiv_rn(fabs((double)delta), __dsqrt_rn(c)));
    return delta == 0.0f ? 0.0f : __double2float_rn(q);
}

// The reference chain of events.c:338-361 with the shortcuts above (bit-identical to tstat_reference_chain).
// BIG: the caller guarantees that every nonzero sample of the windows has |x| >= 2^-60, so that the window sums
// (nonzero multiples of 2^-83) and sums of squares (>= 2^-120) are outside the guarded range of div_w<float>.
template <int W, bool BIG = false>
__device__ __forceinline__ float tstat_fast(double sum1, double ssq1, double sum2d, double ssq2d) {
    const float sum2 = __double2float_rn(sum2d);
    const float ssq2 = __double2float_rn(ssq2d);
    const float mean1 = __double2float_rn(div_w<W>(sum1));
    const float mean2 = div_w<W, !BIG>(sum2);
    const float m1sq = __fmul_rn(mean1, mean1);
    const float m2sq = __fmul_rn(mean2, mean2);
    const float v2 = div_w<W, !BIG>(ssq2);
    double acc = div_w<W>(ssq1);
    acc = __dsub_rn(acc, (double)m1sq);
    acc = __dadd_rn(acc, (double)v2);
    acc = __dsub_rn(acc, (double)m2sq);
    const float cv = fmaxf(__double2float_rn(acc), FLT_MIN);
    const float delta = __fsub_rn(mean2, mean1);
    float scaled = div_w<W, false>(cv);              // cv >= FLT_MIN > 0
    if (cv < 1.0e-36f) scaled = __fdiv_rn(cv, (float)W);  // below 2^-119: outside the validated range of the shortcut
    return tstat_tail(delta, scaled);
}

// Event statistics from the two prefix-sum differences (src/events.c:457-473).
__device__ __forceinline__ void event_stats(double dsum, double dssq, uint32_t len, float* mean,
                                            float* stdv) {
    const float lenf = (float)len;
    const float m = __fdiv_rn(__double2float_rn(dsum), lenf);
    const float ex2 = __fdiv_rn(__double2float_rn(dssq), lenf);
    const float var = __fsub_rn(ex2, __fmul_rn(m, m));
    *mean = m;
    *stdv = __fsqrt_rn(fmaxf(var, 0.0f));
}

// State of one peak detector between samples (src/events.c:269-281).
struct DetState {
    uint32_t masked_to;
    int32_t  peak_pos;   // -1 = none
    float    peak_value;
    int32_t  valid;
};

__device__ __forceinline__ void det_reset(DetState& d) {
    d.masked_to = 0; d.peak_pos = -1; d.peak_value = FLT_MAX; d.valid = 0;
}

// index of the read containing flat position p: largest r with read_off[r] <= p
__device__ __forceinline__ uint32_t find_read(const uint64_t* __restrict__ off, uint32_t n, uint64_t p) {
    uint32_t lo = 0, hi = n;  // invariant: off[lo] <= p < off[hi] (off[n] = span)
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (off[mid] <= p) lo = mid; else hi = mid;
    }
    return lo;
}

}  // namespace sgpu
