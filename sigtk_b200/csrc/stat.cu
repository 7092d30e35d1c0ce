// stat.cu -- the per-read statistics of `sigtk stat` (reference src/cfunc.c:126-159, src/stat.h:17-73).
//
// The reference accumulates in FLOAT, sequentially: the result depends on the
// summation order, so a tree reduction cannot be bit-exact.
// stat_moments_kernel therefore replays the float recurrence in order ("exact
// emulation"), one warp per read with the chains on two lanes.  Medians are order-free: a two-level
// radix selection on the int16 samples, one CTA per read; pa_median follows
// from raw_median because the pA map is monotone in raw.
#include "kernels.cuh"

namespace sgpu {

// One WARP per read. The additions of a float accumulator form one dependent chain (4 cycles each): nothing can
// shorten it, but everything around it can run beside it. All lanes load 32 consecutive samples (coalesced) and
// compute the addends -- (float)raw and pA in the first pass, the squared deviations in the second -- into shared
// memory; lanes 0 and 1 then run the two chains (raw / pA) side by side through the same instructions, while the
// next tile's samples are already in flight.
__global__ void __launch_bounds__(128) stat_moments_kernel(DevBatch b, float* __restrict__ out) {
    __shared__ float add_all[4][2][32];
    float (*add)[32] = add_all[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = warp; r < b.n_reads; r += n_warps) {
        const int16_t* __restrict__ raw = b.samples + b.read_off[r];
        const int n = (int)b.read_len[r];  // stat.h takes `int n`
        const float off = b.offset[r], unit = b.unit[r];
        const float nf = (float)n;
        float mean = 0.0f;  // lane 0: raw, lane 1: pA
        for (int pass = 0; pass < 2; pass++) {
            // pass 0: meani16 / meanf: sum += x[i]; pass 1: stdvi16 / stdvf: sum += (x[i]-m)*(x[i]-m)
            const float mean_r = __shfl_sync(0xffffffffu, mean, 0), mean_p = __shfl_sync(0xffffffffu, mean, 1);
            float acc = 0.0f;
            int16_t nxt = lane < n ? raw[lane] : (int16_t)0;
            for (int t0 = 0; t0 < n; t0 += 32) {
                const int16_t v = nxt;
                nxt = t0 + 32 + lane < n ? raw[t0 + 32 + lane] : (int16_t)0;
                float ar = (float)v, ap = pa_of(v, off, unit);
                if (pass) {
                    const float dr = __fsub_rn(ar, mean_r), dp = __fsub_rn(ap, mean_p);
                    ar = __fmul_rn(dr, dr);
                    ap = __fmul_rn(dp, dp);
                }
                __syncwarp();  // the previous tile's chains are done with the shared tile
                add[0][lane] = ar;
                add[1][lane] = ap;
                __syncwarp();
                if (lane < 2) {
                    const float* a = add[lane];
                    const int cnt = n - t0 < 32 ? n - t0 : 32;
                    if (cnt == 32) {
#pragma unroll
                        for (int k = 0; k < 32; k++) acc = __fadd_rn(acc, a[k]);
                    } else {
                        for (int k = 0; k < cnt; k++) acc = __fadd_rn(acc, a[k]);
                    }
                }
            }
            if (pass == 0) {
                mean = __fdiv_rn(acc, nf);
                if (lane < 2) out[(size_t)r * 6 + lane] = mean;
            } else if (lane < 2) {
                out[(size_t)r * 6 + 2 + lane] = __fsqrt_rn(__fdiv_rn(acc, nf));
            }
        }
    }
}

// element of rank `rank` (0-based, ascending) of one read's int16 samples; all threads of the CTA call it
__device__ int select_rank_i16(const int16_t* __restrict__ raw, uint32_t n, uint32_t rank, uint32_t* hist,
                               uint32_t* sh) {
    // level 1: high byte of the order-preserving key (raw + 32768)
    for (int k = threadIdx.x; k < 256; k += blockDim.x) hist[k] = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&hist[((uint32_t)(raw[i] + 32768)) >> 8], 1u);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t seen = 0, k = 0;
        for (; k < 256; k++) {
            if (seen + hist[k] > rank) break;
            seen += hist[k];
        }
        sh[0] = k;
        sh[1] = rank - seen;
    }
    __syncthreads();
    const uint32_t hi = sh[0], rank2 = sh[1];
    __syncthreads();
    // level 2: low byte among the samples whose high byte matched
    for (int k = threadIdx.x; k < 256; k += blockDim.x) hist[k] = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t key = (uint32_t)(raw[i] + 32768);
        if ((key >> 8) == hi) atomicAdd(&hist[key & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t seen = 0, k = 0;
        for (; k < 256; k++) {
            if (seen + hist[k] > rank2) break;
            seen += hist[k];
        }
        sh[2] = k;
    }
    __syncthreads();
    const int med = (int)((hi << 8) | sh[2]) - 32768;
    __syncthreads();
    return med;
}

__global__ void __launch_bounds__(256) stat_median_kernel(DevBatch b, float* __restrict__ out) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t sh[4];
    for (uint32_t r = blockIdx.x; r < b.n_reads; r += gridDim.x) {
        const int16_t* __restrict__ raw = b.samples + b.read_off[r];
        const uint32_t n = b.read_len[r];
        if (n == 0) continue;
        const uint32_t rank = (uint32_t)((int)n / 2);  // ks_ksmall(n, copy, n/2), stat.h:60,70
        const int med_r = select_rank_i16(raw, n, rank, hist, sh);
        const float off = b.offset[r], unit = b.unit[r];
        // pA is non-decreasing in raw for unit >= 0 and non-increasing for unit < 0
        int med_for_pa = med_r;
        if (unit < 0.0f) med_for_pa = select_rank_i16(raw, n, n - 1 - rank, hist, sh);
        if (threadIdx.x == 0) {
            float* o = out + (size_t)r * 6;
            o[4] = (float)med_r;
            o[5] = pa_of((int16_t)med_for_pa, off, unit);
        }
    }
}

int launch_stat(const DevBatch& b, float* stat6, int sm_count, cudaStream_t st) {
    if (b.n_reads == 0) return 0;
    int g1 = (int)((b.n_reads + 3) / 4);  // one warp per read
    if (g1 > sm_count * 16) g1 = sm_count * 16;
    stat_moments_kernel<<<g1, 128, 0, st>>>(b, stat6);
    int g2 = (int)b.n_reads;
    if (g2 > sm_count * 8) g2 = sm_count * 8;
    stat_median_kernel<<<g2, 256, 0, st>>>(b, stat6);
    return 2;
}

}  // namespace sgpu
