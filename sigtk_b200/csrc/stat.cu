// stat.cu -- the per-read statistics of `sigtk stat` (reference src/cfunc.c:126-159, src/stat.h:17-73).
//
// The reference accumulates in FLOAT, sequentially: the result depends on the
// summation order, so a tree reduction cannot be bit-exact.
// stat_moments_kernel therefore reproduces the float recurrence exactly ("exact
// emulation"), one warp per read: runs of 32 values are summarised in parallel by
// what they do to the accumulator for either parity of its last bit (see below).  Medians are order-free: a two-level
// radix selection on the int16 samples, one CTA per read; pa_median follows
// from raw_median because the pA map is monotone in raw.
#include "kernels.cuh"
#include "float_order.cuh"

namespace sgpu {

// One WARP per read; per pass and superblock of 1024 samples all lanes load the samples (coalesced) and compute the
// addends -- (float)raw and pA in the first pass (meani16 / meanf: sum += x[i]), the squared deviations in the
// second (stdvi16 / stdvf: sum += (x[i]-m)*(x[i]-m)) -- into shared memory; chain_superblock adds them up in the
// reference's order. Positions past the end of the read hold +0, which no float sum notices.
// JNN = true: ONE channel, the samples clamped to [0, 1200] (rm_outlier, jnn.c:58-75), i.e. meanf / stdvf of the
// signal jnn_core thresholds (jnn.c:181-185); out = [n_reads][2] (mean, stdv).
template <bool JNN>
__global__ void __launch_bounds__(128, 4) stat_moments_kernel(DevBatch b, float* __restrict__ out) {
    __shared__ float add_all[4][JNN ? 1 : 2][32 * SB_STRIDE];
    __shared__ int sums_all[4][96];
    float (*add)[32 * SB_STRIDE] = add_all[threadIdx.x >> 5];
    int* sums = sums_all[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = warp; r < b.n_reads; r += n_warps) {
        const int16_t* __restrict__ raw = b.samples + b.read_off[r];
        const int n = (int)b.read_len[r];  // stat.h takes `int n`
        const float off = JNN ? 0.0f : b.offset[r], unit = JNN ? 0.0f : b.unit[r];
        const float nf = (float)n;
        float mean_r = 0.0f, mean_p = 0.0f;
        // 128-bit loads: lane L takes the words q*32 + L of a superblock (4 per lane, 512 contiguous bytes per
        // instruction); the next superblock's are in flight while this one is added up (the kernel is bound by the
        // latency of these loads: one warp per read, profiles/r01_jnn_moments_ncu.md)
        const uint4* __restrict__ src = reinterpret_cast<const uint4*>(raw);
        const int n_words = (n + 7) >> 3;
        uint4 cur[4], nxt[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int w = q * 32 + lane;
            cur[q] = w < n_words ? __ldg(src + w) : make_uint4(0, 0, 0, 0);
        }
        for (int pass = 0; pass < 2; pass++) {
            float acc_r = 0.0f, acc_p = 0.0f;
            for (int t0 = 0; t0 < n; t0 += SB) {
                __syncwarp();  // the previous superblock's readers are done
                {
                    const int tn = t0 + SB < n ? t0 + SB : 0;  // after the last superblock: the first one again (pass 2)
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int w = (tn >> 3) + q * 32 + lane;
                        nxt[q] = w < n_words ? __ldg(src + w) : make_uint4(0, 0, 0, 0);
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t wd[4] = {cur[q].x, cur[q].y, cur[q].z, cur[q].w};
#pragma unroll
                    for (int h = 0; h < 8; h++) {
                        const int j = q * 256 + lane * 8 + h;  // sample of the superblock: tile j/32, column j%32
                        const int i = t0 + j;
                        float ar = 0.0f, ap = 0.0f;
                        if (i < n) {
                            const int16_t v = (int16_t)(wd[h >> 1] >> ((h & 1) * 16));
                            if (JNN) {
                                ar = (float)min(max((int)v, 0), 1200);
                            } else {
                                ar = (float)v;
                                ap = pa_of(v, off, unit);
                            }
                            if (pass) {
                                const float dr = __fsub_rn(ar, mean_r), dp = __fsub_rn(ap, mean_p);
                                ar = __fmul_rn(dr, dr);
                                ap = __fmul_rn(dp, dp);
                            }
                        }
                        const int at = (j >> 5) * SB_STRIDE + (j & 31);
                        add[0][at] = ar;
                        if (!JNN) add[JNN ? 0 : 1][at] = ap;
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; q++) cur[q] = nxt[q];
                __syncwarp();
                const int ntiles = (min(SB, n - t0) + 31) >> 5;
                acc_r = chain_superblock(add[0], ntiles, acc_r, lane, sums, t0 == 0);
                if (!JNN) acc_p = chain_superblock(add[JNN ? 0 : 1], ntiles, acc_p, lane, sums, t0 == 0);
            }
            if (pass == 0) {
                mean_r = __fdiv_rn(acc_r, nf);
                mean_p = __fdiv_rn(acc_p, nf);
                if (lane == 0) {
                    if (JNN) out[(size_t)r * 2] = mean_r;
                    else { out[(size_t)r * 6] = mean_r; out[(size_t)r * 6 + 1] = mean_p; }
                }
            } else if (lane == 0) {
                if (JNN) out[(size_t)r * 2 + 1] = __fsqrt_rn(__fdiv_rn(acc_r, nf));
                else {
                    out[(size_t)r * 6 + 2] = __fsqrt_rn(__fdiv_rn(acc_r, nf));
                    out[(size_t)r * 6 + 3] = __fsqrt_rn(__fdiv_rn(acc_p, nf));
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) stat_median_kernel(DevBatch b, float* __restrict__ out) {
    __shared__ uint32_t hist[MED_BINS];
    __shared__ uint32_t part[256];
    __shared__ uint32_t sh[4];
    for (uint32_t r = blockIdx.x; r < b.n_reads; r += gridDim.x) {
        const int16_t* __restrict__ raw = b.samples + b.read_off[r];
        const uint32_t n = b.read_len[r];
        if (n == 0) {  // an empty record: defined values instead of the previous batch's
            if (threadIdx.x == 0) { out[(size_t)r * 6 + 4] = 0.0f; out[(size_t)r * 6 + 5] = 0.0f; }
            continue;
        }
        const uint32_t rank = (uint32_t)((int)n / 2);  // ks_ksmall(n, copy, n/2), stat.h:60,70
        const int med_r = select_rank_i16<true>(raw, n, rank, hist, part, sh);
        const float off = b.offset[r], unit = b.unit[r];
        // pA is non-decreasing in raw for unit >= 0 and non-increasing for unit < 0
        int med_for_pa = med_r;
        if (unit < 0.0f) med_for_pa = select_rank_i16<true>(raw, n, n - 1 - rank, hist, part, sh);
        if (threadIdx.x == 0) {
            float* o = out + (size_t)r * 6;
            o[4] = (float)med_r;
            o[5] = pa_of((int16_t)med_for_pa, off, unit);
        }
    }
}

int launch_stat_moments(const DevBatch& b, float* stat6, int sm_count, cudaStream_t st) {
    if (b.n_reads == 0) return 0;
    int g1 = (int)((b.n_reads + 3) / 4);  // one warp per read
    if (g1 > sm_count * 16) g1 = sm_count * 16;
    stat_moments_kernel<false><<<g1, 128, 0, st>>>(b, stat6);
    return 1;
}

int launch_stat_median(const DevBatch& b, float* stat6, int sm_count, cudaStream_t st) {
    if (b.n_reads == 0) return 0;
    int g2 = (int)b.n_reads;
    if (g2 > sm_count * 8) g2 = sm_count * 8;
    stat_median_kernel<<<g2, 256, 0, st>>>(b, stat6);
    return 1;
}

// mean and standard deviation (float, in the reference's summation order) of the samples clamped to [0, 1200]:
// what jnn_core derives its band from (jnn.c:58-75, 181-185). moments2 = [n_reads][2].
int launch_jnn_moments(const DevBatch& b, float* moments2, int sm_count, cudaStream_t st) {
    if (b.n_reads == 0) return 0;
    int g1 = (int)((b.n_reads + 3) / 4);  // one warp per read
    if (g1 > sm_count * 16) g1 = sm_count * 16;
    stat_moments_kernel<true><<<g1, 128, 0, st>>>(b, moments2);
    return 1;
}

}  // namespace sgpu
