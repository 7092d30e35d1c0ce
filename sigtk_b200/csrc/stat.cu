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

namespace sgpu {

// ---- sequential float accumulation, in parallel ----------------------------------------------------------------------
// The reference adds in float, one value after the other: s <- fl(s + v). While s stays in one binade [2^e, 2^(e+1))
// every s is an integer multiple S of g = 2^(e-23), and for v >= 0
//     fl(s + v) = g * RNE(S + v/g) = g * (S + a + c),   v/g = a + f,  a = floor, 0 <= f < 1,
//     c = [f > 1/2], and on a tie (f == 1/2) c = parity of (S + a):
// the only thing the step needs to know about the running sum is the PARITY of S. So a run of values is summarised by
// two integers -- the total increment for an even and for an odd incoming S -- computed without knowing s, and runs
// compose. One warp takes 1024 values: every lane summarises 32 consecutive ones (both parities), a shuffle scan
// composes the 32 summaries (lane 0 used to walk them one by one: a quarter of the kernel's time). A tile that would carry S to 2^24 (the binade ends there), holds a negative value or meets s <= 0 is
// added value by value; the summaries after it are recomputed for the new binade. Bit-identical to the sequential
// loop by construction (tests: every stat of every parity test and of the fuzz reads against the oracle).
constexpr int SB = 1024;            // values per superblock
constexpr int SB_STRIDE = 33;       // shared-memory row stride of a 32-value tile (conflict-free both ways)

__device__ __forceinline__ float chain_superblock(const float* __restrict__ A, int ntiles, float s, int lane,
                                                  int* __restrict__ sums, bool first) {
    int tile = 0;
    while (tile < ntiles) {  // (uniform across the warp)
        const uint32_t sb = __float_as_uint(s);
        const int e = (int)((sb >> 23) & 0xffu) - 127;
        const bool s_ok = !first && s > 0.0f && e > -100 && e < 100;  // the first superblock starts from 0: binades fly by
        int fail = tile;  // first tile that has to be added value by value
        uint32_t S = (sb & 0x7fffffu) | 0x800000u;
        if (s_ok) {
            const float scale = __uint_as_float((uint32_t)(127 + 23 - e) << 23);
            // every lane summarises one tile: the increment of S for an even (u0) and for an odd (u1) incoming S
            uint32_t u0 = 0u, u1 = 0u;  // lanes outside [tile, ntiles) are the identity
            bool ok = true;
            if (lane >= tile && lane < ntiles) {
                const float* row = A + lane * SB_STRIDE;
                // A tie rounds to even: after the first tie of the run the parity is 0 whatever came in, so the two
                // summaries differ only by the carry of that first tie (c and 1 - c). One chain (incoming S even)
                // plus that difference.
                int inc0 = 0, dif = 0;
                uint32_t par = 0u;
                bool seen = false;
#pragma unroll 8
                for (int k = 0; k < 32; k++) {
                    const float v = row[k];
                    const float x = __fmul_rn(v, scale);          // exact: a power of two
                    ok = ok && ((__float_as_uint(v) >> 31) == 0u) && (x < 16777216.0f);
                    const int a = __float2int_rd(x);
                    const float f = __fsub_rn(x, (float)a);       // exact
                    const bool tie = f == 0.5f;
                    const uint32_t up = f > 0.5f ? 1u : 0u;
                    const uint32_t t = par + (uint32_t)a;
                    const uint32_t c = tie ? (t & 1u) : up;
                    if (tie && !seen) { dif = 1 - 2 * (int)c; seen = true; }
                    inc0 += a + (int)c;
                    par = (t + c) & 1u;
                }
                u0 = (uint32_t)inc0;
                u1 = (uint32_t)(inc0 + dif);
            }
            // the summaries compose (the parity after a tile is the parity of S + its increment): an inclusive scan
            // over the lanes gives the increment from tile `tile` through every tile for either incoming parity.
            // (32-bit wrap-around can only happen after the first tile that ends the binade, which is all we need.)
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t b0 = __shfl_up_sync(0xffffffffu, u0, d), b1 = __shfl_up_sync(0xffffffffu, u1, d);
                if (lane >= d) {
                    const uint32_t n0 = b0 + ((b0 & 1u) ? u1 : u0);
                    const uint32_t n1 = b1 + ((b1 & 1u) ? u0 : u1);
                    u0 = n0; u1 = n1;
                }
            }
            const uint32_t S_after = S + ((S & 1u) ? u1 : u0);
            const uint32_t bad = __ballot_sync(0xffffffffu, !ok || S_after >= 0x1000000u);
            fail = bad ? __ffs(bad) - 1 : ntiles;  // the first tile that has to be added value by value
            if (fail > tile) S = __shfl_sync(0xffffffffu, S_after, fail - 1);
            s = __uint_as_float(((uint32_t)(e + 127) << 23) | (S & 0x7fffffu));
            __syncwarp();
        }
        if (fail < ntiles) {
            if (lane == 0) {
                const float* row = A + fail * SB_STRIDE;
#pragma unroll
                for (int k = 0; k < 32; k++) s = __fadd_rn(s, row[k]);
            }
            s = __shfl_sync(0xffffffffu, s, 0);
            tile = fail + 1;
        } else {
            tile = ntiles;
        }
    }
    return s;
}

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

// element of rank `rank` (0-based, ascending) of one read's int16 samples; all threads of the CTA call it.
// Two levels over the order-preserving key raw + 32768: the upper 12 bits (4096 bins: the ~1,000 ADC units a signal
// spans spread over ~60 bins, so the shared-memory atomics of a warp rarely meet; with 256 bins they met on 2-4 bins),
// then the lower 4 bits among the samples of the selected bin. 128-bit loads (the read starts 16-byte aligned).
constexpr int MED_BINS = 4096;
__device__ int select_rank_i16(const int16_t* __restrict__ raw, uint32_t n, uint32_t rank, uint32_t* hist,
                               uint32_t* part, uint32_t* sh) {
    const uint4* __restrict__ src = reinterpret_cast<const uint4*>(raw);
    const uint32_t n_words = (n + 7u) >> 3;
    for (int k = threadIdx.x; k < MED_BINS; k += blockDim.x) hist[k] = 0;
    __syncthreads();
    for (uint32_t w = threadIdx.x; w < n_words; w += blockDim.x) {
        const uint4 q = __ldg(src + w);
        const uint32_t wd[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int h = 0; h < 8; h++) {
            const uint32_t key = (uint32_t)((int)(int16_t)(wd[h >> 1] >> ((h & 1) * 16)) + 32768);
            if (w * 8u + h < n) atomicAdd(&hist[key >> 4], 1u);
        }
    }
    __syncthreads();
    {   // 256 threads x 16 bins, then one thread over the 256 partial sums and the 16 bins of the group
        uint32_t sum = 0;
        for (int k = 0; k < MED_BINS / 256; k++) sum += hist[threadIdx.x * (MED_BINS / 256) + k];
        part[threadIdx.x] = sum;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t seen = 0, g = 0;
        for (; g < 256; g++) {
            if (seen + part[g] > rank) break;
            seen += part[g];
        }
        uint32_t k = g * (MED_BINS / 256);
        for (;; k++) {
            if (seen + hist[k] > rank) break;
            seen += hist[k];
        }
        sh[0] = k;
        sh[1] = rank - seen;
    }
    __syncthreads();
    const uint32_t hi = sh[0], rank2 = sh[1];
    __syncthreads();
    // level 2: the low 4 bits among the samples whose upper 12 bits matched
    if (threadIdx.x < 16) part[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t w = threadIdx.x; w < n_words; w += blockDim.x) {
        const uint4 q = __ldg(src + w);
        const uint32_t wd[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int h = 0; h < 8; h++) {
            const uint32_t key = (uint32_t)((int)(int16_t)(wd[h >> 1] >> ((h & 1) * 16)) + 32768);
            if ((key >> 4) == hi && w * 8u + h < n) atomicAdd(&part[key & 15u], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t seen = 0, k = 0;
        for (; k < 16; k++) {
            if (seen + part[k] > rank2) break;
            seen += part[k];
        }
        sh[2] = k;
    }
    __syncthreads();
    const int med = (int)((hi << 4) | sh[2]) - 32768;
    __syncthreads();
    return med;
}

__global__ void __launch_bounds__(256) stat_median_kernel(DevBatch b, float* __restrict__ out) {
    __shared__ uint32_t hist[MED_BINS];
    __shared__ uint32_t part[256];
    __shared__ uint32_t sh[4];
    for (uint32_t r = blockIdx.x; r < b.n_reads; r += gridDim.x) {
        const int16_t* __restrict__ raw = b.samples + b.read_off[r];
        const uint32_t n = b.read_len[r];
        if (n == 0) continue;
        const uint32_t rank = (uint32_t)((int)n / 2);  // ks_ksmall(n, copy, n/2), stat.h:60,70
        const int med_r = select_rank_i16(raw, n, rank, hist, part, sh);
        const float off = b.offset[r], unit = b.unit[r];
        // pA is non-decreasing in raw for unit >= 0 and non-increasing for unit < 0
        int med_for_pa = med_r;
        if (unit < 0.0f) med_for_pa = select_rank_i16(raw, n, n - 1 - rank, hist, part, sh);
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
