// jnn.cu -- `sigtk jnn`: the stall / homopolymer-stretch segmenter of the reference (src/jnn.c:176-282 jnn_core +
// jnn_raw, parameters jnn.h:24-45 as jnn_print picks them, jnn.c:305-312). SURVEY 8f rank 3 (the jnn half).
//
// Per read:   sig = clamp(raw, 0, 1200)                     (rm_outlier, jnn.c:58-75)
//             band = mean(sig) -/+ 0.75 * stdv(sig)         (float sums in sample order: stat.cu, exact replay)
//             a counter machine over one bit per sample ("inside the band"), emitting (start, end) pairs.
// The machine is inherently sequential inside a read (its counters carry across the whole read: `w` never resets),
// and a step is a handful of integer operations, so one THREAD walks one read: 128-bit loads of 8 samples, the
// in-band test on integers (sig is integer-valued, so `sig < top && sig > bot` is `lo_i <= v <= hi_i` for the
// integers just inside the float band), words of 8 samples that leave the machine idle are skipped at once.
// Reads are independent, so a batch keeps every SM busy as long as it holds a few thousand reads; the walk of one
// very long read is latency bound (documented in DESIGN.md).
#include "kernels.cuh"

namespace sgpu {

struct JnnParams {  // jnn.h:24-45
    int window;
    float stall_len;
    int error, seg_dist, corrector;
    float std_scale;
};

__host__ __device__ inline JnnParams jnn_params(int rna) {
    JnnParams p;
    p.std_scale = 0.75f; p.corrector = 50; p.seg_dist = 50; p.error = 5;
    if (rna) { p.window = 1000; p.stall_len = 1.0f; }   // JNNV1_DRNA_R9_PARAM
    else     { p.window = 150;  p.stall_len = 0.25f; }  // JNNV1_CDNA_R9_PARAM
    return p;
}

// seg: pairs (x, y) of read r at seg[2*(base(r)+k)], base(r) = read_off[r]/32 + r (a read of n samples has at most
// n/38 + 1 segments: the first needs >= window*stall_len >= 37.5 samples, every other >= window).
__global__ void __launch_bounds__(128) jnn_walk_kernel(DevBatch b, const float* __restrict__ moments,
                                                       uint32_t* __restrict__ seg_cnt, int32_t* __restrict__ seg) {
    const JnnParams P = jnn_params(b.rna);
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < b.n_reads; r += gridDim.x * blockDim.x) {
        const int16_t* __restrict__ raw = b.samples + b.read_off[r];
        const int n = (int)b.read_len[r];
        int32_t* out = seg + 2 * ((b.read_off[r] >> 5) + r);
        if (n == 0) { seg_cnt[r] = 0; continue; }
        const float mn = moments[2 * r], sd = moments[2 * r + 1];
        const float band = __fmul_rn(sd, P.std_scale);                        // jnn.c:184-185
        const float top = __fadd_rn(mn, band), bot = __fsub_rn(mn, band);
        // sig is an integer in [0, 1200]: sig < top <=> sig <= ceil(top)-1, sig > bot <=> sig >= floor(bot)+1
        // (NaN band -- never for finite input -- leaves the band empty like the float comparisons do)
        int hi_i = -1, lo_i = 0;
        if (top == top && bot == bot) {
            hi_i = (int)ceilf(fminf(fmaxf(top, -1.0f), 2000.0f)) - 1;
            lo_i = (int)floorf(fminf(fmaxf(bot, -2.0f), 2000.0f)) + 1;
        }
        const float first_min = __fmul_rn((float)P.window, P.stall_len);      // jnn.c:232
        int prev = 0, err = 0, prev_err = 0, c = 0, w = P.corrector, start = 0, n_seg = 0, last_y = 0;
        for (int i0 = 0; i0 < n; i0 += 8) {
            int16_t v[8];
            if (i0 + 8 <= n) {
                *reinterpret_cast<uint4*>(v) = __ldg(reinterpret_cast<const uint4*>(raw + i0));
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] = (i0 + j < n) ? raw[i0 + j] : (int16_t)-32768;
            }
            uint32_t in = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int s = min(max((int)v[j], 0), 1200);
                in |= (uint32_t)(s >= lo_i && s <= hi_i) << j;
            }
            const int m = min(8, n - i0);
            if (m < 8) in &= (1u << m) - 1u;
            if (!prev && in == 0) continue;  // outside the band with no stretch open: nothing moves
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (j >= m) break;
                const int i = i0 + j;
                if ((in >> j) & 1u) {                                          // jnn.c:206-219
                    if (!prev) { start = i; prev = 1; }
                    c++; w++;
                    prev_err = 0;
                    if (c >= P.window && c >= w && !(c % w)) err--;
                } else if (prev) {
                    if (err < P.error) {                                       // jnn.c:222-229
                        c++; err++; prev_err++;
                        if (c >= P.window && c >= w && !(c % w)) err--;
                    } else {
                        if (c >= P.window || (!n_seg && (float)c >= first_min)) {   // jnn.c:230-249
                            const int end = i - prev_err;
                            if (n_seg && start - last_y < P.seg_dist) {
                                out[2 * (n_seg - 1) + 1] = end;
                            } else {
                                out[2 * n_seg] = start;
                                out[2 * n_seg + 1] = end;
                                n_seg++;
                            }
                            last_y = end;
                        }
                        prev = 0; c = 0; err = 0; prev_err = 0;                // jnn.c:246-248, 251-256
                    }
                }
            }
        }
        seg_cnt[r] = (uint32_t)n_seg;
    }
}

uint64_t jnn_seg_capacity(uint64_t max_samples, uint32_t max_reads) { return max_samples / 32 + max_reads + 1; }

// moments: [n_reads][2] scratch; seg_cnt: [n_reads]; seg: [2 * jnn_seg_capacity]
int launch_jnn(const DevBatch& b, float* moments, uint32_t* seg_cnt, int32_t* seg, int sm_count, cudaStream_t st) {
    if (b.n_reads == 0) return 0;
    int n = launch_jnn_moments(b, moments, sm_count, st);
    int g = (int)((b.n_reads + 127) / 128);
    // few reads: spread them over the SMs (32 threads per block keep one warp per scheduler busy at best)
    int threads = 128;
    if (b.n_reads < (uint32_t)sm_count * 128u) { threads = 32; g = (int)((b.n_reads + 31) / 32); }
    if (g > sm_count * 16) g = sm_count * 16;
    jnn_walk_kernel<<<g, threads, 0, st>>>(b, moments, seg_cnt, seg);
    return n + 1;
}

}  // namespace sgpu
