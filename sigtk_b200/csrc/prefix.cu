// prefix.cu -- `sigtk prefix` (reference src/cfunc.c:169-234): the adaptor stretch at the start of a direct-RNA
// read, the poly-A stretch behind it, and the float statistics of both.
//
//   find_adaptor = jnnv2 (jnn.c:99-179) with JNNV2_RNA_R9_ADAPTOR / JNNV2_RNA_RNA004_ADAPTOR (jnn.h:88-102):
//     the raw samples clamped to [0, 1200] (rm_outlier, jnn.c:58-75), their rolling mean over 2,000 samples
//     (rolling_window, jnn.c:20-50: a running FLOAT sum -- of integers <= 1200 * 2000 < 2^24, hence exact in any
//     order -- divided by the window), meanf / stdvf of the rolling means IN SAMPLE ORDER (stat.h:17-44: order
//     dependent, reproduced with the parity-map scan of float_order.cuh), the runs of rolling means below
//     mean - std_scale * stdv merged when closer than 1,500, and the first run of a plausible length;
//   meanf / stdvf / medianf of the adaptor's pA (cfunc.c:178-180);
//   find_polya = jnn_pa + jnn_core (jnn.c:190-279, 352-374) with JNNV1_R9_POLYA on the pA behind the adaptor and the
//     band (adaptor mean + 30) -/+ 20 (cfunc.c:189): a stretch opens at an in-band sample and closes at the 31st
//     out-of-band sample after it (the reference's corrector never acts: it would need run >= 50 + every in-band sample
//     of the record, but run <= the stretch's in-band samples + 30); only the first segment is wanted, extended while
//     the following ones start within 200 samples of the last end;
//   the same three statistics of the poly-A stretch (cfunc.c:202-204).
//
// prefix_walk_kernel: one WARP per read (everything above except the medians); prefix_median_kernel: one CTA per read,
// two radix selections on sub-ranges of the int16 samples (the pA map is monotone in raw).
#include "float_order.cuh"
#include "kernels.cuh"

namespace sgpu {

namespace {

constexpr int PW = 2000;          // jnnv2 window
constexpr int P_MERGE = 1500;     // seg_dist
constexpr int P_HI = 200000;      // hi_thresh
constexpr int A_WINDOW = 250;     // JNNV1_R9_POLYA: window (stall_len 1.0: the same bound for the first segment)
constexpr int A_TOLERATED = 30;   // error
constexpr int A_MERGE = 200;      // seg_dist

__device__ __forceinline__ int clamp_raw(int v) { return min(max(v, 0), 1200); }   // rm_outlier, jnn.c:58-75
__device__ __forceinline__ uint32_t clamp2(uint32_t two) {                          // the same on two packed int16
    return __vmins2(__vmaxs2(two, 0u), 1200u * 0x10001u);
}
static_assert(PW % 8 == 0, "the two ranges of a window difference are loaded with the same alignment");

// mean and standard deviation (float, sample order) of the pA of raw[a .. a + len)
__device__ void range_moments(const int16_t* __restrict__ raw, int a, int len, float off, float unit, float* add,
                              int lane, float* mean_out, float* stdv_out) {
    const float nf = (float)len;
    float mean = 0.0f;
    for (int pass = 0; pass < 2; pass++) {
        float acc = 0.0f;
        for (int t0 = 0; t0 < len; t0 += SB) {
            __syncwarp();
#pragma unroll 4
            for (int q = 0; q < 32; q++) {       // tile q, column lane: sample t0 + 32 q + lane (coalesced)
                const int i = t0 + q * 32 + lane;
                float v = 0.0f;
                if (i < len) {
                    v = pa_of(__ldg(raw + a + i), off, unit);
                    if (pass) { const float d = __fsub_rn(v, mean); v = __fmul_rn(d, d); }
                }
                add[q * SB_STRIDE + lane] = v;
            }
            __syncwarp();
            acc = chain_superblock(add, (min(SB, len - t0) + 31) >> 5, acc, lane, t0 == 0);
        }
        if (pass == 0) { mean = __fdiv_rn(acc, nf); *mean_out = mean; }
        else *stdv_out = __fsqrt_rn(__fdiv_rn(acc, nf));
    }
}

struct RunState {   // jnnv2's scan over the rolling means (jnn.c:125-153) and its pick (jnn.c:155-167)
    int begin, start, end;
    int have, lx, ly;      // the last segment (still open to merging)
    int chosen, cx, cy;
};
__device__ __forceinline__ bool plausible(int a, int b, int lo) { return !(b - a > P_HI) && !(b - a < lo); }
__device__ __forceinline__ void close_run(RunState& s, int lo) {
    if (s.have && s.start - s.ly < P_MERGE) {
        s.ly = s.end;
    } else {
        if (s.have && !s.chosen && plausible(s.lx, s.ly, lo)) { s.chosen = 1; s.cx = s.lx; s.cy = s.ly; }
        s.lx = s.start; s.ly = s.end; s.have = 1;
    }
    s.start = 0; s.end = 0; s.begin = 0;
}

}  // namespace

// pos4[r] = {adaptor x, adaptor y, poly-A x, poly-A y} as jnnv2 / find_polya return them (poly-A relative to the
// adaptor's end; (0,0): no adaptor, (-1,-1): record not longer than the window / no poly-A / DNA);
// st6[r] = mean, stdv of the adaptor's pA, [2] unused here (median: prefix_median_kernel), the same for the poly-A.
__global__ void __launch_bounds__(128, 4) prefix_walk_kernel(DevBatch b, float std_scale, int lo_thresh,
                                                            int32_t* __restrict__ pos4, float* __restrict__ st6) {
    __shared__ float add_all[4][32 * SB_STRIDE];
    float* add = add_all[threadIdx.x >> 5];
    int* dbuf = reinterpret_cast<int*>(add);
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = warp; r < b.n_reads; r += n_warps) {
        const int16_t* __restrict__ raw = b.samples + b.read_off[r];
        const int n = (int)b.read_len[r];
        const float off = b.offset[r], unit = b.unit[r];
        int ax = -1, ay = -1, px = -1, py = -1;
        float st[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
        if (n > PW) {
            const int nt = n - PW;
            const float ntf = (float)nt;
            // the first window's sum (exact integer)
            int s_first = 0;
            for (int w = lane; w < PW / 8; w += 32) {   // 128-bit loads (a read starts 16-byte aligned)
                const uint4 q = __ldg(reinterpret_cast<const uint4*>(raw) + w);
                const uint32_t c4[4] = {clamp2(q.x), clamp2(q.y), clamp2(q.z), clamp2(q.w)};
#pragma unroll
                for (int h = 0; h < 4; h++) s_first += (int)(c4[h] & 0xffffu) + (int)(c4[h] >> 16);
            }
            for (int o = 16; o; o >>= 1) s_first += __shfl_xor_sync(0xffffffffu, s_first, o);
            float mn = 0.0f, floor_ = 0.0f;
            RunState rs = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int pass = 0; pass < 3 && !rs.chosen; pass++) {
                float acc = 0.0f;
                int s_base = s_first;
                for (int t0 = 0; t0 < nt; t0 += SB) {
                    __syncwarp();
                    // d[i] = c[i + PW] - c[i]: the window sum moves by d[i] from position i to i + 1.
                    // Lane L takes the 128-bit words q * 32 + L of both ranges (PW is a multiple of 8: they are
                    // aligned alike; eight loads in flight per lane), two samples per clamp / subtraction
                    {
                        const uint4* __restrict__ lo = reinterpret_cast<const uint4*>(raw + t0);
                        const uint4* __restrict__ hi = reinterpret_cast<const uint4*>(raw + t0 + PW);
                        uint4 a[4], c[4];
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const int wq = q * 32 + lane;
                            const bool in = t0 + wq * 8 < nt;   // (the word's first position; the read is padded to 8 samples)
                            a[q] = in ? __ldg(lo + wq) : make_uint4(0, 0, 0, 0);
                            c[q] = in ? __ldg(hi + wq) : make_uint4(0, 0, 0, 0);
                        }
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const uint32_t av[4] = {a[q].x, a[q].y, a[q].z, a[q].w}, cv[4] = {c[q].x, c[q].y, c[q].z, c[q].w};
#pragma unroll
                            for (int h = 0; h < 4; h++) {
                                const uint32_t d2 = __vsub2(clamp2(cv[h]), clamp2(av[h]));   // two differences in [-1200, 1200]
                                const int j = q * 256 + lane * 8 + 2 * h;                     // position in the superblock
                                const int i = t0 + j;
                                const int d0 = (int)(int16_t)(d2 & 0xffffu), d1 = (int)d2 >> 16;
                                dbuf[(j >> 5) * SB_STRIDE + (j & 31)] = i < nt ? d0 : 0;
                                dbuf[(j >> 5) * SB_STRIDE + (j & 31) + 1] = i + 1 < nt ? d1 : 0;
                            }
                        }
                    }
                    __syncwarp();
                    // lane L owns tile L: its 32 consecutive positions
                    int* row = dbuf + lane * SB_STRIDE;
                    int rowsum = 0;
#pragma unroll 8
                    for (int k = 0; k < 32; k++) rowsum += row[k];
                    int incl = rowsum;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += t;
                    }
                    int s = s_base + incl - rowsum;
                    uint32_t below = 0u, above = 0u;
#pragma unroll 8
                    for (int k = 0; k < 32; k++) {
                        const int i = t0 + lane * 32 + k;
                        const int d = row[k];
                        float v = 0.0f;
                        if (i < nt) {
                            // rolling_window, jnn.c:41,46: (float)s / 2000, by the reciprocal and one residual
                            // correction (equal to the IEEE quotient for every integer s <= 2^24:
                            // oracle/proofs/div_window_check.c)
                            const float sf = (float)s;
                            const float q0 = __fmul_rn(sf, 1.0f / (float)PW);
                            const float t = __fmaf_rn(__fmaf_rn(-(float)PW, q0, sf), 1.0f / (float)PW, q0);
                            if (pass == 0) v = t;
                            else if (pass == 1) { const float dv = __fsub_rn(t, mn); v = __fmul_rn(dv, dv); }
                            else { below |= (t < floor_ ? 1u : 0u) << k; above |= (t > floor_ ? 1u : 0u) << k; }
                        }
                        reinterpret_cast<float*>(row)[k] = v;
                        s += d;
                    }
                    s_base += __shfl_sync(0xffffffffu, incl, 31);
                    __syncwarp();
                    const int ntiles = (min(SB, nt - t0) + 31) >> 5;
                    if (pass < 2) {
                        acc = chain_superblock(add, ntiles, acc, lane, t0 == 0);
                    } else {
                        // the run logic, every lane the same (uniform): tile after tile, transition after transition
                        for (int tl = 0; tl < ntiles && !rs.chosen; tl++) {
                            const uint32_t bm = __shfl_sync(0xffffffffu, below, tl), am = __shfl_sync(0xffffffffu, above, tl);
                            const int p0 = t0 + tl * 32;
                            int cur = 0;
                            while (cur < 32) {
                                const uint32_t rest = 0xffffffffu << cur;
                                if (!rs.begin) {
                                    const uint32_t m = bm & rest;
                                    if (!m) break;
                                    const int k = __ffs(m) - 1;
                                    rs.start = p0 + k; rs.begin = 1;
                                    cur = k + 1;
                                } else {
                                    const uint32_t ma = am & rest;
                                    const int ka = ma ? __ffs(ma) - 1 : 32;
                                    const uint32_t mb = bm & rest & (ka < 32 ? ((1u << ka) - 1u) : 0xffffffffu);
                                    if (mb) rs.end = p0 + 31 - __clz(mb);
                                    if (ka == 32) break;
                                    close_run(rs, lo_thresh);
                                    cur = ka + 1;
                                }
                            }
                        }
                    }
                }
                if (pass == 0) mn = __fdiv_rn(acc, ntf);
                else if (pass == 1) floor_ = __fsub_rn(mn, __fmul_rn(__fsqrt_rn(__fdiv_rn(acc, ntf)), std_scale));
            }
            if (!rs.chosen && rs.have && plausible(rs.lx, rs.ly, lo_thresh)) { rs.chosen = 1; rs.cx = rs.lx; rs.cy = rs.ly; }
            ax = 0; ay = 0;
            if (rs.chosen) { ax = rs.cx + PW / 2 - 1; ay = rs.cy + PW / 2 - 1; }
            if (ay > 0) {
                range_moments(raw, ax, ay - ax, off, unit, add, lane, &st[0], &st[1]);
                if (b.rna) {
                    // ---- poly-A: the counter machine of jnn_core over one bit per sample (in the band or not) ----
                    const float top = __fadd_rn(__fadd_rn(st[0], 30.0f), 20.0f), bot = __fsub_rn(__fadd_rn(st[0], 30.0f), 20.0f);
                    const int m = n - ay;
                    int open = 0, first = 0, bad = 0, tail = 0, n_seg = 0, sx0 = 0, sy0 = 0, last_y = 0;
                    for (int t0 = 0; t0 < m && n_seg < 2; t0 += 32) {
                        const int i = t0 + lane;
                        bool in = false;
                        if (i < m) {
                            float v = pa_of(__ldg(raw + ay + i), off, unit);
                            v = v > 1200.0f ? 1200.0f : v < 0.0f ? 0.0f : v;   // rm_outlierf, jnn.c:79-95
                            in = v < top && v > bot;
                        }
                        const uint32_t w = __ballot_sync(0xffffffffu, in);
                        const uint32_t valid = m - t0 >= 32 ? 0xffffffffu : ((1u << (m - t0)) - 1u);
                        int cur = 0;
                        while (cur < 32 && ((valid >> cur) & 1u) && n_seg < 2) {
                            const uint32_t rest = (0xffffffffu << cur) & valid;
                            if (!open) {
                                const uint32_t mi = w & rest;
                                if (!mi) break;
                                const int k = __ffs(mi) - 1;
                                open = 1; first = t0 + k; bad = 0; tail = 0;
                                cur = k + 1;
                            } else {
                                const uint32_t z = ~w & rest;                  // out-of-band samples from cur on
                                const int nz = __popc(z);
                                if (bad + nz <= A_TOLERATED) {                 // the stretch survives this word
                                    bad += nz;
                                    const uint32_t ones = w & rest;
                                    const int hi = 31 - __clz(valid);          // last valid bit
                                    tail = ones ? hi - (31 - __clz(ones)) : tail + (hi + 1 - cur);
                                    break;
                                }
                                const int k = (int)__fns(z, 0u, A_TOLERATED - bad + 1);   // the 31st outlier of the stretch
                                const uint32_t before = w & rest & ((1u << k) - 1u);      // in-band samples in [cur, k)
                                const int tl = before ? k - 1 - (31 - __clz(before)) : tail + (k - cur);
                                const int iclose = t0 + k;
                                if (iclose - first >= A_WINDOW) {              // run = every sample of the stretch
                                    const int stop = iclose - tl;
                                    if (n_seg && first - last_y < A_MERGE) { if (n_seg == 1) sy0 = stop; }
                                    else { if (n_seg == 0) { sx0 = first; sy0 = stop; } n_seg++; }
                                    last_y = stop;
                                }
                                open = 0; bad = 0; tail = 0;
                                cur = k + 1;
                            }
                        }
                    }
                    if (n_seg > 0) { px = sx0; py = sy0; }
                    if (py > 0) range_moments(raw, ay + px, py - px, off, unit, add, lane, &st[3], &st[4]);
                }
            }
        }
        if (lane == 0) {
            int32_t* p = pos4 + (size_t)r * 4;
            p[0] = ax; p[1] = ay; p[2] = px; p[3] = py;
            float* o = st6 + (size_t)r * 6;
            o[0] = st[0]; o[1] = st[1]; o[3] = st[3]; o[4] = st[4];
        }
    }
}

// medianf of the two stretches: the sample of rank len/2 (stat.h:56-64, ks_ksmall) among the raw values of the
// stretch, mapped to pA (monotone in raw; rank from the other end for a negative unit)
__global__ void __launch_bounds__(256) prefix_median_kernel(DevBatch b, const int32_t* __restrict__ pos4,
                                                            float* __restrict__ st6) {
    __shared__ uint32_t hist[MED_BINS];
    __shared__ uint32_t part[256];
    __shared__ uint32_t sh[4];
    for (uint32_t r = blockIdx.x; r < b.n_reads; r += gridDim.x) {
        const int16_t* __restrict__ raw = b.samples + b.read_off[r];
        const int32_t* p = pos4 + (size_t)r * 4;
        const int ax = p[0], ay = p[1], px = p[2], py = p[3];
        const float off = b.offset[r], unit = b.unit[r];
        float med[2] = {0.0f, 0.0f};
        for (int k = 0; k < 2; k++) {
            const int a = k ? ay + px : ax, len = k ? py - px : ay - ax;
            const bool have = k ? (ay > 0 && py > 0) : ay > 0;   // (uniform across the CTA)
            if (!have || len <= 0) continue;
            uint32_t rank = (uint32_t)(len / 2);
            if (unit < 0.0f) rank = (uint32_t)len - 1u - rank;
            const int v = select_rank_i16<false>(raw + a, (uint32_t)len, rank, hist, part, sh);
            med[k] = pa_of((int16_t)v, off, unit);
        }
        if (threadIdx.x == 0) { st6[(size_t)r * 6 + 2] = med[0]; st6[(size_t)r * 6 + 5] = med[1]; }
    }
}

int launch_prefix(const DevBatch& b, int rna004, int32_t* pos4, float* st6, int sm_count, cudaStream_t st) {
    if (b.n_reads == 0) return 0;
    int g1 = (int)((b.n_reads + 3) / 4);  // one warp per read
    if (g1 > sm_count * 16) g1 = sm_count * 16;
    // JNNV2_RNA_R9_ADAPTOR / JNNV2_RNA_RNA004_ADAPTOR (jnn.h:88-102)
    prefix_walk_kernel<<<g1, 128, 0, st>>>(b, rna004 ? 0.7f : 0.5f, rna004 ? 500 : 2000, pos4, st6);
    int g2 = (int)b.n_reads;
    if (g2 > sm_count * 8) g2 = sm_count * 8;
    prefix_median_kernel<<<g2, 256, 0, st>>>(b, pos4, st6);
    return 2;
}

}  // namespace sgpu
