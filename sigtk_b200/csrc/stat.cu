// stat.cu -- the per-read statistics of `sigtk stat` (reference src/cfunc.c:126-159, src/stat.h:17-73).
//
// The reference accumulates in FLOAT, sequentially: the result depends on the
// summation order, so a tree reduction cannot be bit-exact.
// stat_moments_kernel therefore replays the float recurrence in order, one
// thread per read ("exact emulation").  Medians are order-free: a two-level
// radix selection on the int16 samples, one CTA per read; pa_median follows
// from raw_median because the pA map is monotone in raw.
#include "kernels.cuh"

namespace sgpu {

__global__ void __launch_bounds__(128) stat_moments_kernel(DevBatch b, float* __restrict__ out) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < b.n_reads; r += gridDim.x * blockDim.x) {
        const int16_t* __restrict__ raw = b.samples + b.read_off[r];
        const int n = (int)b.read_len[r];  // stat.h takes `int n`
        const float off = b.offset[r], unit = b.unit[r];
        const float nf = (float)n;
        float acc_r = 0.0f, acc_p = 0.0f;  // meani16 / meanf: sum += x[i]
        for (int i = 0; i < n; i++) {
            const int16_t v = raw[i];
            acc_r = __fadd_rn(acc_r, (float)v);
            acc_p = __fadd_rn(acc_p, pa_of(v, off, unit));
        }
        const float mean_r = __fdiv_rn(acc_r, nf), mean_p = __fdiv_rn(acc_p, nf);
        float dev_r = 0.0f, dev_p = 0.0f;  // stdvi16 / stdvf: sum += (x[i]-m)*(x[i]-m)
        for (int i = 0; i < n; i++) {
            const int16_t v = raw[i];
            const float dr = __fsub_rn((float)v, mean_r);
            const float dp = __fsub_rn(pa_of(v, off, unit), mean_p);
            dev_r = __fadd_rn(dev_r, __fmul_rn(dr, dr));
            dev_p = __fadd_rn(dev_p, __fmul_rn(dp, dp));
        }
        float* o = out + (size_t)r * 6;
        o[0] = mean_r;
        o[1] = mean_p;
        o[2] = __fsqrt_rn(__fdiv_rn(dev_r, nf));
        o[3] = __fsqrt_rn(__fdiv_rn(dev_p, nf));
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
    int g1 = (int)((b.n_reads + 127) / 128);
    if (g1 > sm_count * 16) g1 = sm_count * 16;
    stat_moments_kernel<<<g1, 128, 0, st>>>(b, stat6);
    int g2 = (int)b.n_reads;
    if (g2 > sm_count * 8) g2 = sm_count * 8;
    stat_median_kernel<<<g2, 256, 0, st>>>(b, stat6);
    return 2;
}

}  // namespace sgpu
