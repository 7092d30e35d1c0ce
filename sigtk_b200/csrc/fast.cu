// fast.cu -- the tiled fast path of the sigtk B200 hot path (sm_100a).
//
//   detect_tiles_kernel   per tile of the flat sample array: TMA bulk load of the int16 samples into shared
//                         memory (double buffered, mbarrier completion), pA conversion fused into the load
//                         (misc.c:15-32; pA optionally stored), segmented FP64 inclusive prefix sums of x and x*x
//                         (events.c:293-303), both window t-statistics in one pass (events.c:315-364), and the
//                         dual peak detector (events.c:371-443) run as 64-sample chunks that start W samples early
//                         from a cold state; chunk boundary states are compared and mismatching chunks re-run from
//                         the true state.  Output: one bit per sample (event starts) + boundary states per tile +
//                         per-read exact-sum witness.
//   verify_tiles_kernel   compares the detector state across tile boundaries; a mismatch routes the read to the
//                         sequential-order kernels (generic.cu) and is counted in fixups[read].
//   build_seq_list_kernel evaluates the exact-sum witness per read and compacts the list of reads that need the
//                         sequential-order kernels.
//   count_tile_bits_kernel / read_event_offsets_kernel   ranks of event starts (ev_off).
//   emit_tiles_kernel     second pass: event table (create_events/create_event, events.c:457-504) from the
//                         event-start bits and freshly recomputed prefix sums.
//
// Why the results are bit-identical to the reference although the sums are formed in another order: the
// reference's S[i], Q[i] (double) are exact (no rounding happened) whenever every value is a multiple of 2^k and
// the sum of magnitudes stays below 2^(k+53); then ANY grouping of the same additions is exact too, and every
// window / event difference S[b]-S[a] equals the exact sum of the samples in [a,b).  The witness checks that
// sufficient condition per read from min|x| and max|x|; reads that fail it (rare: samples within ~1 pA of zero
// in long reads) are recomputed in the reference's own order by generic.cu.
#include "kernels.cuh"

namespace sgpu {

namespace {

constexpr int T = FAST_TILE;            // core samples per tile
constexpr int NT = 288;                 // threads of detect_tiles_kernel
constexpr int REG = NT * 8;             // samples staged per tile (core + halos), 8 per thread
constexpr int L = 64;                   // detector chunk length
constexpr int NCH = T / L;              // chunks per tile = one warp
constexpr int SEG_MAX = 64;             // reads intersecting one staged region
constexpr int NONE = INT_MIN;
static_assert(NCH == 32, "one detector chunk per lane of one warp");

template <int RNA>
struct Geo {
    static constexpr int w1 = RNA ? 7 : 3;
    static constexpr int w2 = RNA ? 14 : 6;
    static constexpr int W = RNA ? 128 : 32;   // detector warm-up before a chunk
    static constexpr int R = RNA ? 64 : 32;    // detector run-out after a chunk
    static constexpr int HL = ((W + w2 + 1) + 7) / 8 * 8;
    static constexpr int HR = ((R + w2) + 7) / 8 * 8;
    static_assert(HL + T + HR <= REG, "region too small");
};

struct Seg {           // one read intersecting the staged region, in region coordinates
    int u0;            // region index of the read's first sample (may be negative)
    uint32_t len;
    uint32_t read;
    float off, unit;
};

__device__ __forceinline__ int pad8(int u) { return u + (u >> 3); }     // doubles: 8-sample groups, stride 9
__device__ __forceinline__ int pad64(int u) { return u + (u >> 6); }    // floats: 64-sample chunks, stride 65

// ---- mbarrier / bulk-copy (TMA) helpers ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- shared helpers --------------------------------------------------------------------------------------------
// List the reads that intersect flat range [lo, hi) (executed by one thread).
// When more than SEG_MAX reads intersect the range (reads of a few dozen samples) the tile is not handled by the
// fast path: `overflow` is raised and, if seq_flag is given, the reads beyond the list are routed to the
// sequential-order kernels here (the caller flags the listed ones).
__device__ int collect_segments(const DevBatch& b, long long lo, long long hi, long long region_start, Seg* segs,
                                int* overflow, uint32_t* seq_flag, uint32_t r) {
    if (hi <= 0 || b.n_reads == 0) return 0;
    if (lo < 0) lo = 0;
    int n = 0;
    for (; r < b.n_reads; r++) {
        const long long s = (long long)b.read_off[r];
        if (s >= hi) break;
        const uint32_t len = b.read_len[r];
        if (len == 0 || s + (long long)len <= lo) continue;
        if (n == SEG_MAX) {
            *overflow = 1;
            if (!seq_flag) break;
            seq_flag[r] = 1u;
            continue;
        }
        segs[n].u0 = (int)(s - region_start);
        segs[n].len = len;
        segs[n].read = r;
        segs[n].off = b.offset[r];
        segs[n].unit = b.unit[r];
        n++;
    }
    return n;
}

// index of the segment that owns the 8-sample group starting at region index u (or -1: alignment gap / outside)
__device__ __forceinline__ int group_segment(const Seg* segs, int nseg, int u) {
    int lo = 0, hi = nseg;  // last segment with u0 <= u
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (segs[mid].u0 <= u) lo = mid + 1; else hi = mid;
    }
    const int s = lo - 1;
    if (s < 0) return -1;
    return ((long long)u - segs[s].u0 < (long long)segs[s].len) ? s : -1;
}

// Segmented inclusive prefix sums of x and x*x over the staged region, 8 consecutive samples per thread.
// Writes sS/sQ (padded); x outside reads counts as 0; the sums restart at every read start.
template <int NTHREADS>
__device__ __forceinline__ void region_prefix(const float (&x)[8], bool starts_read, double* sS, double* sQ,
                                              double* warp_v, int* warp_f) {
    constexpr int NW = NTHREADS / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double s[8], q[8];
    double as = 0.0, aq = 0.0;
#pragma unroll
    for (int m = 0; m < 8; m++) {
        as = __dadd_rn(as, (double)x[m]);
        aq = __dadd_rn(aq, (double)__fmul_rn(x[m], x[m]));
        s[m] = as;
        q[m] = aq;
    }
    // segmented inclusive scan of the thread totals across the warp
    double vs = as, vq = aq;
    int f = starts_read ? 1 : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double us = __shfl_up_sync(0xffffffffu, vs, o);
        const double uq = __shfl_up_sync(0xffffffffu, vq, o);
        const int uf = __shfl_up_sync(0xffffffffu, f, o);
        if (lane >= o) {
            if (!f) { vs = __dadd_rn(vs, us); vq = __dadd_rn(vq, uq); }
            f |= uf;
        }
    }
    if (lane == 31) { warp_v[2 * wid] = vs; warp_v[2 * wid + 1] = vq; warp_f[wid] = f; }
    double es = __shfl_up_sync(0xffffffffu, vs, 1), eq = __shfl_up_sync(0xffffffffu, vq, 1);
    int ef = __shfl_up_sync(0xffffffffu, f, 1);
    if (lane == 0) { es = 0.0; eq = 0.0; ef = 0; }
    __syncthreads();
    double cs = 0.0, cq = 0.0;  // carry from the warps before this one
    for (int k = 0; k < NW; k++) {
        if (k < wid) {
            if (warp_f[k]) { cs = warp_v[2 * k]; cq = warp_v[2 * k + 1]; }
            else { cs = __dadd_rn(cs, warp_v[2 * k]); cq = __dadd_rn(cq, warp_v[2 * k + 1]); }
        }
    }
    double bs, bq;  // exclusive prefix of this thread
    if (starts_read) { bs = 0.0; bq = 0.0; }
    else if (ef) { bs = es; bq = eq; }
    else { bs = __dadd_rn(cs, es); bq = __dadd_rn(cq, eq); }
    const int base = pad8(threadIdx.x * 8);
#pragma unroll
    for (int m = 0; m < 8; m++) {
        sS[base + m] = __dadd_rn(bs, s[m]);
        sQ[base + m] = __dadd_rn(bq, q[m]);
    }
}

__device__ __forceinline__ void unpack8(const int4& raw, int (&v)[8]) {
    const int w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        v[2 * k] = (int)(int16_t)(w[k] & 0xffff);
        v[2 * k + 1] = w[k] >> 16;
    }
}

// ---- the dual peak detector on region coordinates --------------------------------------------------------------
struct Det {
    int mt;      // masked_to
    int pp;      // peak_pos or NONE
    float pv;    // peak_value
    int valid;
};
__device__ __forceinline__ void det_set(Det& d, int masked_to) { d.mt = masked_to; d.pp = NONE; d.pv = FLT_MAX; d.valid = 0; }

// one sample of one detector (events.c:387-437); returns the emitted peak position or NONE
template <bool SHORT>
__device__ __forceinline__ int det_step(Det& d, Det& lng, int u, float cur, float thr, int w, int w_short, float h) {
    if (d.mt >= u) return NONE;
    if (d.pp == NONE) {
        if (cur < d.pv) d.pv = cur;
        else if (__fsub_rn(cur, d.pv) > h) { d.pv = cur; d.pp = u; }
        return NONE;
    }
    if (cur > d.pv) { d.pv = cur; d.pp = u; }
    if (SHORT && d.pv > thr) { lng.mt = d.pp + w_short; lng.pp = NONE; lng.pv = FLT_MAX; lng.valid = 0; }
    if (__fsub_rn(d.pv, cur) > h && d.pv > thr) d.valid = 1;
    if (d.valid && (u - d.pp) > w / 2) {
        const int out = d.pp;
        d.pp = NONE; d.pv = cur; d.valid = 0;
        return out;
    }
    return NONE;
}

struct DetPair { Det s, l; };

// canonical form of the pair at boundary b (state before sample b is processed), positions relative to b
struct Canon { int v[8]; };
__device__ __forceinline__ Canon canon(const DetPair& p, int b) {
    Canon c;
    c.v[0] = p.s.mt >= b ? p.s.mt - b : -1;
    c.v[1] = p.s.pp == NONE ? NONE : p.s.pp - b;
    c.v[2] = __float_as_int(p.s.pv);
    c.v[3] = p.s.valid;
    c.v[4] = p.l.mt >= b ? p.l.mt - b : -1;
    c.v[5] = p.l.pp == NONE ? NONE : p.l.pp - b;
    c.v[6] = __float_as_int(p.l.pv);
    c.v[7] = p.l.valid;
    return c;
}
__device__ __forceinline__ bool canon_eq(const Canon& a, const Canon& b) {
    bool e = true;
#pragma unroll
    for (int k = 0; k < 8; k++) e = e && (a.v[k] == b.v[k]);
    return e;
}

template <int RNA>
struct Walker {
    const float* t1;
    const float* t2;
    const signed char* grp;
    const Seg* segs;
    int lo, hi;       // region range of the peaks this chunk owns (hi = lo + 64)

    // run samples [a, b) through both detectors; emissions owned by the chunk set bit (pos - lo) of `mask`
    __device__ __noinline__ void run(DetPair& p, int a, int b, unsigned long long& mask) const {
        using G = Geo<RNA>;
        const DetParams prm = det_params(RNA);
        int sidx = -2, su0 = 0, send = 0;
        for (int u = a; u < b; u++) {
            if (sidx == -2 || (u & 7) == 0) {
                sidx = grp[u >> 3];
                if (sidx >= 0) { su0 = segs[sidx].u0; send = (int)min((long long)su0 + (long long)segs[sidx].len, (long long)REG); }
            }
            if (sidx < 0 || u >= send) continue;           // alignment gap
            if (u == su0) { det_set(p.s, u); det_set(p.l, u); }  // first sample of a read: initial state (516-536)
            const float c1 = t1[pad64(u)], c2 = t2[pad64(u)];
            const int e1 = det_step<true>(p.s, p.l, u, c1, prm.thr1, G::w1, G::w1, prm.height);
            const int e2 = det_step<false>(p.l, p.l, u, c2, prm.thr2, G::w2, G::w1, prm.height);
            if (e1 != NONE && e1 >= lo && e1 < hi) mask |= 1ull << (e1 - lo);
            if (e2 != NONE && e2 >= lo && e2 < hi) mask |= 1ull << (e2 - lo);
        }
    }

    // can the detector still emit a peak that this chunk owns?
    __device__ __forceinline__ bool pending(const Det& d, float thr) const {
        return d.pp != NONE && d.pp >= lo && d.pp < hi && (d.valid || d.pv > thr);
    }

    // continue past the chunk end until the owned pending peaks are resolved; false if the cap was hit
    __device__ __noinline__ bool run_out(DetPair p, int from, unsigned long long& mask) const {
        using G = Geo<RNA>;
        const DetParams prm = det_params(RNA);
        int u = from;
        const int cap = from + G::R;
        while (pending(p.s, prm.thr1) || pending(p.l, prm.thr2)) {
            if (u >= cap) return false;
            const int sidx = grp[u >> 3];
            if (sidx < 0) return true;                              // the read ended: pending peaks are dropped
            const int su0 = segs[sidx].u0;
            if (u == su0 || (long long)u - su0 >= (long long)segs[sidx].len) return true;
            run(p, u, u + 1, mask);
            u++;
        }
        return true;
    }
};

// Both detectors for one sample that lies inside a read and is not its first sample, without branches
// (same transitions as det_step: events.c:387-437). Peaks the chunk owns set bit (pos - lo) of `mask`.
template <int RNA>
__device__ __forceinline__ void step_pair(DetPair& p, int u, float c1, float c2, int lo, unsigned long long& mask) {
    using G = Geo<RNA>;
    constexpr float thr1 = RNA ? 2.5f : 1.4f, thr2 = 9.0f, h = RNA ? 1.0f : 0.2f;
    {
        Det& d = p.s;
        const bool act = d.mt < u, none = d.pp == NONE;
        const bool lt = c1 < d.pv, gt = c1 > d.pv;
        const bool rise = __fsub_rn(c1, d.pv) > h;
        const float pv2 = gt ? c1 : d.pv;
        const int pp2 = gt ? u : d.pp;
        const bool big = pv2 > thr1;
        const bool valid2 = (d.valid != 0) | (big & (__fsub_rn(pv2, c1) > h));
        const bool emit = valid2 & ((int)((unsigned)u - (unsigned)pp2) > G::w1 / 2);
        const bool in1 = act & none, in2 = act & !none;
        const bool maskl = in2 & big;  // the short detector dominates the long one (414-422)
        p.l.mt = maskl ? pp2 + G::w1 : p.l.mt;
        p.l.pp = maskl ? NONE : p.l.pp;
        p.l.pv = maskl ? FLT_MAX : p.l.pv;
        p.l.valid = maskl ? 0 : p.l.valid;
        const unsigned k = (unsigned)((in2 & emit) ? pp2 - lo : -1);
        mask |= (k < 64u) ? (1ull << k) : 0ull;
        d.pv = in1 ? ((lt | rise) ? c1 : d.pv) : (in2 ? (emit ? c1 : pv2) : d.pv);
        d.pp = in1 ? ((!lt & rise) ? u : NONE) : (in2 ? (emit ? NONE : pp2) : d.pp);
        d.valid = in2 ? ((valid2 & !emit) ? 1 : 0) : d.valid;
    }
    {
        Det& d = p.l;
        const bool act = d.mt < u, none = d.pp == NONE;
        const bool lt = c2 < d.pv, gt = c2 > d.pv;
        const bool rise = __fsub_rn(c2, d.pv) > h;
        const float pv2 = gt ? c2 : d.pv;
        const int pp2 = gt ? u : d.pp;
        const bool big = pv2 > thr2;
        const bool valid2 = (d.valid != 0) | (big & (__fsub_rn(pv2, c2) > h));
        const bool emit = valid2 & ((int)((unsigned)u - (unsigned)pp2) > G::w2 / 2);
        const bool in1 = act & none, in2 = act & !none;
        const unsigned k = (unsigned)((in2 & emit) ? pp2 - lo : -1);
        mask |= (k < 64u) ? (1ull << k) : 0ull;
        d.pv = in1 ? ((lt | rise) ? c2 : d.pv) : (in2 ? (emit ? c2 : pv2) : d.pv);
        d.pp = in1 ? ((!lt & rise) ? u : NONE) : (in2 ? (emit ? NONE : pp2) : d.pp);
        d.valid = in2 ? ((valid2 & !emit) ? 1 : 0) : d.valid;
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
struct DetectSmem {
    alignas(16) int16_t raw[2][REG];
    alignas(8) uint64_t bar[2];
    double sS[REG + REG / 8 + 8];
    double sQ[REG + REG / 8 + 8];
    float t1[REG + REG / 64 + 4];
    float t2[REG + REG / 64 + 4];
    Seg segs[SEG_MAX];
    double warp_v[2 * (NT / 32)];
    int warp_f[NT / 32];
    uint32_t bits[T / 32];
    uint32_t wmin[SEG_MAX], wmax[SEG_MAX];
    signed char grp[REG / 8];
    int nseg, overflow, bad;
};

template <int RNA>
__global__ void __launch_bounds__(NT, 3) detect_tiles_kernel(DevBatch b, uint32_t n_tiles, float* __restrict__ pa_out,
                                                          uint32_t* __restrict__ bitmap, int* __restrict__ st_begin,
                                                          int* __restrict__ st_end, uint32_t* __restrict__ wit_min,
                                                          uint32_t* __restrict__ wit_max,
                                                          uint32_t* __restrict__ seq_flag,
                                                          const uint32_t* __restrict__ tile_read0) {
    using G = Geo<RNA>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DetectSmem& sm = *reinterpret_cast<DetectSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long span = (long long)b.span;

    auto region_start_of = [&](uint32_t t) { return (long long)t * T - G::HL; };
    auto issue_load = [&](uint32_t t, int buf) {  // one thread: bulk copy of the valid part of the region
        const long long rs = region_start_of(t);
        const long long lo = rs < 0 ? 0 : rs;
        long long hi = rs + REG;
        if (hi > span) hi = span;
        if (hi > lo) {
            const uint32_t bytes = (uint32_t)(hi - lo) * 2u;
            mbar_expect_tx(&sm.bar[buf], bytes);
            bulk_g2s(&sm.raw[buf][lo - rs], b.samples + lo, bytes, &sm.bar[buf]);
        } else {
            mbar_expect_tx(&sm.bar[buf], 0);
        }
    };

    if (tid == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
        if (blockIdx.x < n_tiles) issue_load(blockIdx.x, 0);
    }
    __syncthreads();

    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
        const int buf = it & 1;
        const long long rs = region_start_of(tile);
        // prefetch the next tile of this CTA into the other buffer (its previous contents were consumed before
        // the barriers at the end of the previous iteration)
        if (tid == 0) {
            fence_proxy_async();
            const uint32_t nxt = tile + gridDim.x;
            if (nxt < n_tiles) issue_load(nxt, buf ^ 1);
            int ovf = 0;
            sm.nseg = collect_segments(b, rs, rs + REG, rs, sm.segs, &ovf, seq_flag, tile_read0[tile]);
            sm.overflow = ovf;
            sm.bad = 0;
        }
        if (tid < T / 32) sm.bits[tid] = 0u;
        if (tid < SEG_MAX) { sm.wmin[tid] = 0xffffffffu; sm.wmax[tid] = 0u; }
        __syncthreads();
        const int nseg = sm.nseg;
        mbar_wait(&sm.bar[buf], (it >> 1) & 1);

        // ---- phase B: pA + segmented prefix sums -------------------------------------------------------------
        {
            const int u0 = tid * 8;
            const int sidx = group_segment(sm.segs, nseg, u0);
            sm.grp[tid] = (signed char)sidx;
            float x[8];
            bool starts = false;
            if (sidx >= 0) {
                const Seg sg = sm.segs[sidx];
                starts = (sg.u0 == u0);
                const int4 rawv = *reinterpret_cast<const int4*>(&sm.raw[buf][u0]);
                int v[8];
                unpack8(rawv, v);
                const uint32_t left = sg.len - (uint32_t)(u0 - sg.u0);  // samples of the read from u0 on
                uint32_t mn = 0xffffffffu, mx = 0u;
#pragma unroll
                for (int m = 0; m < 8; m++) {
                    const float xv = __fmul_rn(__fadd_rn((float)v[m], sg.off), sg.unit);
                    x[m] = ((uint32_t)m < left) ? xv : 0.0f;
                    const uint32_t a = __float_as_uint(x[m]) & 0x7fffffffu;
                    mx = max(mx, a);
                    mn = min(mn, a ? a : 0xffffffffu);
                }
                const bool core = (u0 >= G::HL && u0 < G::HL + T);
                if (core) {
                    atomicMin(&sm.wmin[sidx], mn);
                    atomicMax(&sm.wmax[sidx], mx);
                    if (pa_out) {
                        float4* dst = reinterpret_cast<float4*>(pa_out + (rs + u0));
                        dst[0] = make_float4(x[0], x[1], x[2], x[3]);
                        dst[1] = make_float4(x[4], x[5], x[6], x[7]);
                    }
                }
            } else {
#pragma unroll
                for (int m = 0; m < 8; m++) x[m] = 0.0f;
            }
            region_prefix<NT>(x, starts, sm.sS, sm.sQ, sm.warp_v, sm.warp_f);
        }
        __syncthreads();

        // ---- phase C: both t-statistics for every position the detector will look at ---------------------------
        for (int u = (G::HL - G::W) + tid; u < G::HL + T + G::R; u += NT) {
            float r1 = 0.0f, r2 = 0.0f;
            const int sidx = sm.grp[u >> 3];
            if (sidx >= 0) {
                const int su0 = sm.segs[sidx].u0;
                const uint32_t n = sm.segs[sidx].len;
                const uint32_t i = (uint32_t)(u - su0);
                if (i < n) {
                    const double s_i = sm.sS[pad8(u - 1)], q_i = sm.sQ[pad8(u - 1)];  // i >= w >= 2 where used
                    if (n >= 2u * G::w1 && i >= (uint32_t)G::w1 && i + G::w1 <= n) {
                        const bool first = (i == (uint32_t)G::w1);
                        const double sl = first ? 0.0 : sm.sS[pad8(u - G::w1 - 1)];
                        const double ql = first ? 0.0 : sm.sQ[pad8(u - G::w1 - 1)];
                        r1 = tstat_fast<G::w1>(__dsub_rn(s_i, sl), __dsub_rn(q_i, ql),
                                               __dsub_rn(sm.sS[pad8(u + G::w1 - 1)], s_i),
                                               __dsub_rn(sm.sQ[pad8(u + G::w1 - 1)], q_i));
                    }
                    if (n >= 2u * G::w2 && i >= (uint32_t)G::w2 && i + G::w2 <= n) {
                        const bool first = (i == (uint32_t)G::w2);
                        const double sl = first ? 0.0 : sm.sS[pad8(u - G::w2 - 1)];
                        const double ql = first ? 0.0 : sm.sQ[pad8(u - G::w2 - 1)];
                        r2 = tstat_fast<G::w2>(__dsub_rn(s_i, sl), __dsub_rn(q_i, ql),
                                               __dsub_rn(sm.sS[pad8(u + G::w2 - 1)], s_i),
                                               __dsub_rn(sm.sQ[pad8(u + G::w2 - 1)], q_i));
                    }
                }
            }
            sm.t1[pad64(u)] = r1;
            sm.t2[pad64(u)] = r2;
        }
        // witness: one global atomic pair per read of the tile
        if (tid < nseg && sm.wmax[tid] | (sm.wmin[tid] != 0xffffffffu)) {
            atomicMin(&wit_min[sm.segs[tid].read], sm.wmin[tid]);
            atomicMax(&wit_max[sm.segs[tid].read], sm.wmax[tid]);
        }
        __syncthreads();

        // ---- phase D: the peak detector, one 64-sample chunk per lane of warp 0 ----------------------------------
        if (wid == 0) {
            const int cs = G::HL + lane * L, ce = cs + L;
            const int wa = cs - G::W, wz = ce + G::R;   // everything this lane may look at
            Walker<RNA> wk{sm.t1, sm.t2, sm.grp, sm.segs, cs, ce};
            const DetParams prm = det_params(RNA);
            DetPair p;
            det_set(p.s, wa - 1);  // cold start: wa is the first sample processed
            det_set(p.l, wa - 1);
            unsigned long long mask = 0ull;
            Canon begin, end;
            bool ok = true;
            // common case: [wa, wz) lies inside one read and does not contain its first sample
            bool simple = false;
            {
                const int sa = sm.grp[wa >> 3];
                if (sa >= 0) simple = sm.segs[sa].u0 < wa && (long long)sm.segs[sa].u0 + (long long)sm.segs[sa].len >= wz;
            }
            if (simple) {
#pragma unroll 8
                for (int u = wa; u < cs; u++) step_pair<RNA>(p, u, sm.t1[pad64(u)], sm.t2[pad64(u)], cs, mask);
                begin = canon(p, cs);
#pragma unroll 8
                for (int u = cs; u < ce; u++) step_pair<RNA>(p, u, sm.t1[pad64(u)], sm.t2[pad64(u)], cs, mask);
                end = canon(p, ce);
                DetPair r = p;
                int u = ce;
                while (wk.pending(r.s, prm.thr1) || wk.pending(r.l, prm.thr2)) {
                    if (u >= wz) { ok = false; break; }
                    step_pair<RNA>(r, u, sm.t1[pad64(u)], sm.t2[pad64(u)], cs, mask);
                    u++;
                }
            } else {
                unsigned long long ignore = 0ull;
                wk.run(p, wa, cs, ignore);
                begin = canon(p, cs);
                wk.run(p, cs, ce, mask);
                end = canon(p, ce);
                ok = wk.run_out(p, ce, mask);
            }
            const Canon tile_begin = begin;  // lane 0: speculative state at the tile start (verified across tiles)
            // compare with the previous chunk's end state; re-run mismatching chunks from the true state
            for (int round = 0; round < NCH; round++) {
                Canon prev;
#pragma unroll
                for (int k = 0; k < 8; k++) prev.v[k] = __shfl_up_sync(0xffffffffu, end.v[k], 1);
                const bool mism = lane > 0 && !canon_eq(prev, begin);
                if (!__any_sync(0xffffffffu, mism)) break;
                if (mism) {
                    // rebuild the true state at cs from the canonical form
                    p.s.mt = prev.v[0] >= 0 ? prev.v[0] + cs : cs - 1;
                    p.s.pp = prev.v[1] == NONE ? NONE : prev.v[1] + cs;
                    p.s.pv = __int_as_float(prev.v[2]);
                    p.s.valid = prev.v[3];
                    p.l.mt = prev.v[4] >= 0 ? prev.v[4] + cs : cs - 1;
                    p.l.pp = prev.v[5] == NONE ? NONE : prev.v[5] + cs;
                    p.l.pv = __int_as_float(prev.v[6]);
                    p.l.valid = prev.v[7];
                    begin = prev;
                    mask = 0ull;
                    wk.run(p, cs, ce, mask);
                    end = canon(p, ce);
                    ok = wk.run_out(p, ce, mask);
                }
                __syncwarp();
            }
            sm.bits[2 * lane] = (uint32_t)mask;  // this chunk owns exactly these two words
            sm.bits[2 * lane + 1] = (uint32_t)(mask >> 32);
            if (!ok) {  // run-out cap hit: let the sequential-order kernels do the read that contains ce-1
                const int sidx = sm.grp[(ce - 1) >> 3];
                if (sidx >= 0) seq_flag[sm.segs[sidx].read] = 1u;
            }
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 8; k++) st_begin[(size_t)tile * 8 + k] = tile_begin.v[k];
            }
            if (lane == 31) {
#pragma unroll
                for (int k = 0; k < 8; k++) st_end[(size_t)tile * 8 + k] = end.v[k];
            }
        }
        __syncthreads();
        // event 0 of every read starts at its first sample (events.c:490-497)
        if (tid < nseg) {
            const int c = sm.segs[tid].u0 - G::HL;
            if (c >= 0 && c < T) atomicOr(&sm.bits[c >> 5], 1u << (c & 31));
            if (sm.overflow) seq_flag[sm.segs[tid].read] = 1u;
        }
        __syncthreads();
        if (tid < T / 32) bitmap[(size_t)tile * (T / 32) + tid] = sm.bits[tid];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Tile t (t >= 1) started its first chunk from a speculative state; it must equal the end state of tile t-1
// whenever the boundary lies strictly inside a read.
__global__ void __launch_bounds__(256) verify_tiles_kernel(DevBatch b, uint32_t n_tiles, const int* __restrict__ st_begin,
                                                           const int* __restrict__ st_end,
                                                           uint32_t* __restrict__ seq_flag,
                                                           uint32_t* __restrict__ fixups) {
    for (uint32_t t = 1 + blockIdx.x * blockDim.x + threadIdx.x; t < n_tiles; t += gridDim.x * blockDim.x) {
        const uint64_t p = (uint64_t)t * T;
        if (p >= b.span) continue;
        const uint32_t r = find_read(b.read_off, b.n_reads, p);
        const uint64_t s = b.read_off[r];
        if (p <= s || p >= s + b.read_len[r]) continue;  // read start or alignment gap: nothing carried over
        bool same = true;
        for (int k = 0; k < 8; k++) same = same && (st_begin[(size_t)t * 8 + k] == st_end[(size_t)(t - 1) * 8 + k]);
        if (!same) {
            seq_flag[r] = 1u;
            atomicAdd(&fixups[r], 1u);
        }
    }
}

__device__ __forceinline__ bool sums_exact(uint32_t e_min, uint32_t e_max, uint32_t log2n) {
    // every value is a multiple of 2^(e_min-150) and the sum of magnitudes is below 2^(e_max+1-127+log2n)
    return e_min != 0u && e_max < 255u && (e_max + 1u + log2n <= e_min + 30u);
}

__global__ void __launch_bounds__(256) build_seq_list_kernel(DevBatch b, const uint32_t* __restrict__ wit_min,
                                                             const uint32_t* __restrict__ wit_max,
                                                             uint32_t* __restrict__ seq_flag, int force_all,
                                                             uint32_t* __restrict__ seq_list,
                                                             uint64_t* __restrict__ seq_sbase,
                                                             uint32_t* __restrict__ seq_count,
                                                             unsigned long long* __restrict__ cursor, uint64_t gen_cap,
                                                             int* __restrict__ status,
                                                             unsigned long long* __restrict__ counters) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < b.n_reads; r += gridDim.x * blockDim.x) {
        const uint32_t n = b.read_len[r];
        bool flag = force_all || seq_flag[r] != 0u;
        if (!flag && n > 0) {
            const uint32_t mn = wit_min[r], mx = wit_max[r];
            if (mx != 0u) {  // not all zero
                const uint32_t log2n = n > 1 ? 32u - __clz(n - 1) : 0u;
                const float fmn = __uint_as_float(mn), fmx = __uint_as_float(mx);
                const uint32_t qmn = __float_as_uint(__fmul_rn(fmn, fmn)), qmx = __float_as_uint(__fmul_rn(fmx, fmx));
                flag = !(sums_exact(mn >> 23, mx >> 23, log2n) && sums_exact(qmn >> 23, qmx >> 23, log2n));
            }
        }
        if (n == 0) flag = false;
        seq_flag[r] = flag ? 1u : 0u;
        if (flag) {
            const uint64_t need = ((uint64_t)n + 7u) & ~7ull;
            const unsigned long long base = atomicAdd(cursor, (unsigned long long)need);
            if (base + need > gen_cap) { atomicExch(status, SGPU_DEV_E_SCRATCH); continue; }
            const uint32_t k = atomicAdd(seq_count, 1u);
            seq_list[k] = r;
            seq_sbase[k] = base;
            atomicAdd(&counters[1], 1ull);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) count_tile_bits_kernel(const uint32_t* __restrict__ bitmap, uint32_t n_tiles,
                                                              uint32_t* __restrict__ tile_cnt) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t t = warp; t < n_tiles; t += n_warps) {
        const uint2 v = reinterpret_cast<const uint2*>(bitmap + (size_t)t * (T / 32))[lane];
        uint32_t c = __popc(v.x) + __popc(v.y);
        for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) tile_cnt[t] = c;
    }
}

// number of event starts before flat position p
__device__ __forceinline__ uint64_t rank_before(const uint32_t* __restrict__ bitmap, const uint64_t* __restrict__ tile_base,
                                                uint64_t p) {
    const uint64_t t = p / T;
    const uint32_t u = (uint32_t)(p - t * T);
    uint64_t k = tile_base[t];
    const uint32_t* w = bitmap + t * (T / 32);
    for (uint32_t j = 0; j < (u >> 5); j++) k += __popc(w[j]);
    if (u & 31) k += __popc(w[u >> 5] & ((1u << (u & 31)) - 1u));
    return k;
}

__global__ void __launch_bounds__(256) read_event_offsets_kernel(DevBatch b, const uint32_t* __restrict__ bitmap,
                                                                 const uint64_t* __restrict__ tile_base,
                                                                 uint32_t n_tiles, uint64_t* __restrict__ ev_off) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r <= b.n_reads; r += gridDim.x * blockDim.x) {
        if (r == b.n_reads) { ev_off[r] = tile_base[n_tiles]; continue; }
        const uint64_t p = b.read_off[r];
        ev_off[r] = (p >= b.span) ? tile_base[n_tiles] : rank_before(bitmap, tile_base, p);
    }
}

// ---------------------------------------------------------------------------------------------------------------
constexpr int ENT = 256;  // threads of emit_tiles_kernel (8 samples each = one tile)

struct EmitSmem {
    double sS[T + T / 8 + 8];
    double sQ[T + T / 8 + 8];
    Seg segs[SEG_MAX];
    double warp_v[2 * (ENT / 32)];
    int warp_f[ENT / 32];
    uint32_t bits[T / 32];
    uint32_t excl[T / 32];
    signed char grp[T / 8];
    int nseg, overflow;
    int spill_u;            // tile index of the start of the event that runs past the tile end, or -1
    unsigned long long spill_k;
};

__global__ void __launch_bounds__(ENT) emit_tiles_kernel(DevBatch b, uint32_t n_tiles, const uint32_t* __restrict__ bitmap,
                                                         const uint64_t* __restrict__ tile_base, uint64_t ev_cap,
                                                         uint32_t* __restrict__ ev_start, float* __restrict__ ev_mean,
                                                         float* __restrict__ ev_stdv, int* __restrict__ status,
                                                         const uint32_t* __restrict__ tile_read0) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EmitSmem& sm = *reinterpret_cast<EmitSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long ts = (long long)tile * T;
        if (tid == 0) {
            int ovf = 0;
            sm.nseg = collect_segments(b, ts, ts + T, ts, sm.segs, &ovf, nullptr, tile_read0[tile]);
            sm.overflow = ovf;  // such reads were routed to the sequential-order kernels by detect_tiles_kernel
            sm.spill_u = -1;
        }
        if (tid < T / 32) sm.bits[tid] = bitmap[(size_t)tile * (T / 32) + tid];
        __syncthreads();
        const int nseg = sm.nseg;
        const uint64_t base = tile_base[tile];
        if (tid < 32) {  // exclusive popcount scan of the 64 words
            const uint32_t c0 = __popc(sm.bits[2 * lane]), c1 = __popc(sm.bits[2 * lane + 1]);
            uint32_t inc = c0 + c1;
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            sm.excl[2 * lane] = inc - c0 - c1;
            sm.excl[2 * lane + 1] = inc - c1;
        }
        const int u0 = tid * 8;
        const int sidx = group_segment(sm.segs, nseg, u0);
        sm.grp[tid] = (signed char)sidx;
        {
            float x[8];
            bool starts = false;
            if (sidx >= 0 && ts + u0 < (long long)b.span) {
                const Seg sg = sm.segs[sidx];
                starts = (sg.u0 == u0);
                const int4 rawv = __ldg(reinterpret_cast<const int4*>(b.samples + ts + u0));
                int v[8];
                unpack8(rawv, v);
                const uint32_t left = sg.len - (uint32_t)(u0 - sg.u0);
#pragma unroll
                for (int m = 0; m < 8; m++)
                    x[m] = ((uint32_t)m < left) ? __fmul_rn(__fadd_rn((float)v[m], sg.off), sg.unit) : 0.0f;
            } else {
#pragma unroll
                for (int m = 0; m < 8; m++) x[m] = 0.0f;
            }
            region_prefix<ENT>(x, starts, sm.sS, sm.sQ, sm.warp_v, sm.warp_f);
        }
        __syncthreads();
        // every event start bit in this thread's 8 samples
        if (sidx >= 0) {
            const Seg sg = sm.segs[sidx];
            uint32_t byte = (sm.bits[u0 >> 5] >> (u0 & 31)) & 0xffu;
            while (byte) {
                const int m = __ffs(byte) - 1;
                byte &= byte - 1;
                const int u = u0 + m;
                const uint64_t k = base + sm.excl[u >> 5] + __popc(sm.bits[u >> 5] & ((1u << (u & 31)) - 1u));
                // end of the event: the next start bit, or the end of the read
                const long long read_end = (long long)sg.u0 + (long long)sg.len;  // tile coordinates
                int e = -1;
                {
                    uint32_t w = sm.bits[u >> 5] & ~((2u << (u & 31)) - 1u);
                    int j = u >> 5;
                    while (true) {
                        if (w) { e = (j << 5) + __ffs(w) - 1; break; }
                        if (++j >= T / 32) break;
                        w = sm.bits[j];
                    }
                }
                if (k >= ev_cap) { atomicExch(status, SGPU_DEV_E_EVCAP); continue; }
                ev_start[k] = (uint32_t)(u - sg.u0);
                long long end = (e >= 0) ? e : (long long)T + 1;  // T+1: beyond the tile
                if (end > read_end) end = read_end;
                if (end <= T) {
                    const int eu = (int)end;
                    const bool at0 = (u == sg.u0) || (u == 0);  // nothing of this read before u inside the tile
                    const double s0 = at0 ? 0.0 : sm.sS[pad8(u - 1)];
                    const double q0 = at0 ? 0.0 : sm.sQ[pad8(u - 1)];
                    float mean, stdv;
                    event_stats(__dsub_rn(sm.sS[pad8(eu - 1)], s0), __dsub_rn(sm.sQ[pad8(eu - 1)], q0),
                                (uint32_t)(eu - u), &mean, &stdv);
                    ev_mean[k] = mean;
                    ev_stdv[k] = stdv;
                } else {
                    sm.spill_u = u;  // at most one event per tile runs past the tile end
                    sm.spill_k = k;
                }
            }
        }
        __syncthreads();
        // the event that continues into the following tiles: warp 0 walks the flat array
        if (tid < 32 && sm.spill_u >= 0) {
            const int u = sm.spill_u;
            const int sx = sm.grp[u >> 3];
            const Seg sg = sm.segs[sx];
            const long long read_end = ts + sg.u0 + (long long)sg.len;  // flat
            const bool at0 = (u == sg.u0) || (u == 0);
            double ds = __dsub_rn(sm.sS[pad8(T - 1)], at0 ? 0.0 : sm.sS[pad8(u - 1)]);
            double dq = __dsub_rn(sm.sQ[pad8(T - 1)], at0 ? 0.0 : sm.sQ[pad8(u - 1)]);
            // find the next event start at or after the tile end (bounded by the end of the read)
            long long end = read_end;
            for (long long wbase = (ts + T) >> 5; (wbase << 5) < read_end; wbase += 32) {
                const long long wi = wbase + lane;
                uint32_t w = ((wi << 5) < read_end) ? bitmap[wi] : 0u;
                const uint32_t any = __ballot_sync(0xffffffffu, w != 0u);
                if (any) {
                    const int first = __ffs(any) - 1;
                    const uint32_t fw = __shfl_sync(0xffffffffu, w, first);
                    const long long cand = ((wbase + first) << 5) + __ffs(fw) - 1;
                    if (cand < end) end = cand;
                    break;
                }
            }
            // add the samples [tile end, end): 8 per lane per step
            double as = 0.0, aq = 0.0;
            for (long long p = ts + T + (long long)lane * 8; p < end; p += 256) {
                const int4 rawv = __ldg(reinterpret_cast<const int4*>(b.samples + p));
                int v[8];
                unpack8(rawv, v);
#pragma unroll
                for (int m = 0; m < 8; m++) {
                    if (p + m < end) {
                        const float xv = __fmul_rn(__fadd_rn((float)v[m], sg.off), sg.unit);
                        as = __dadd_rn(as, (double)xv);
                        aq = __dadd_rn(aq, (double)__fmul_rn(xv, xv));
                    }
                }
            }
            for (int o = 16; o; o >>= 1) {
                as = __dadd_rn(as, __shfl_xor_sync(0xffffffffu, as, o));
                aq = __dadd_rn(aq, __shfl_xor_sync(0xffffffffu, aq, o));
            }
            if (lane == 0) {
                float mean, stdv;
                event_stats(__dadd_rn(ds, as), __dadd_rn(dq, aq), (uint32_t)(end - (ts + u)), &mean, &stdv);
                ev_mean[sm.spill_k] = mean;
                ev_stdv[sm.spill_k] = stdv;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) init_reads_kernel(DevBatch b, uint32_t n_tiles, uint32_t* __restrict__ wit_min,
                                                         uint32_t* __restrict__ wit_max, uint32_t* __restrict__ seq_flag,
                                                         uint32_t* __restrict__ fixups, uint32_t* __restrict__ seq_count,
                                                         unsigned long long* __restrict__ cursor,
                                                         uint32_t* __restrict__ tile_read0) {
    const uint32_t n_reads = b.n_reads;
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g == 0) { *seq_count = 0u; *cursor = 0ull; }
    // first read that can intersect the staged region of every tile (the region starts at most 160 samples early)
    for (uint32_t t = g; t < n_tiles; t += gridDim.x * blockDim.x) {
        const long long p = (long long)t * T - 160;
        tile_read0[t] = find_read(b.read_off, n_reads, (uint64_t)(p < 0 ? 0 : p));
    }
    for (uint32_t r = g; r < n_reads; r += gridDim.x * blockDim.x) {
        wit_min[r] = 0xffffffffu;
        wit_max[r] = 0u;
        seq_flag[r] = 0u;
        fixups[r] = 0u;
    }
}

__global__ void __launch_bounds__(256) sum_fixups_kernel(uint32_t n_reads, const uint32_t* __restrict__ fixups,
                                                         unsigned long long* __restrict__ counters) {
    unsigned long long acc = 0;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += gridDim.x * blockDim.x) acc += fixups[r];
    if (acc) atomicAdd(&counters[2], acc);
}

static inline int grid_cap(uint64_t work, int block, int max_blocks) {
    uint64_t g = (work + block - 1) / block;
    if (g < 1) g = 1;
    if (g > (uint64_t)max_blocks) g = max_blocks;
    return (int)g;
}

uint32_t fast_tiles_for(uint64_t span) { return (uint32_t)((span + T - 1) / T); }

int fast_configure() {
    cudaError_t e = cudaFuncSetAttribute(detect_tiles_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(DetectSmem));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(detect_tiles_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(DetectSmem));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(emit_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EmitSmem));
    return e == cudaSuccess ? 0 : -1;
}

int launch_init_reads(const DevBatch& b, Scratch& sc, uint32_t* seq_flag, uint32_t* fixups, int sm_count,
                      cudaStream_t st) {
    const uint32_t n_tiles = fast_tiles_for(b.span);
    init_reads_kernel<<<grid_cap(max(b.n_reads, n_tiles), 256, sm_count * 8), 256, 0, st>>>(
        b, n_tiles, sc.wit_min, sc.wit_max, seq_flag, fixups, sc.seq_count, sc.cursor, sc.tile_read0);
    return 1;
}

int launch_fast_detect(const DevBatch& b, Scratch& sc, float* pa_out, uint32_t* seq_flag, uint32_t* fixups,
                       int sm_count, cudaStream_t st) {
    const uint32_t n_tiles = fast_tiles_for(b.span);
    const int ctas_per_sm = (int)((227u * 1024u) / (sizeof(DetectSmem) + 1024u));
    const int grid = grid_cap(n_tiles, 1, sm_count * (ctas_per_sm > 0 ? ctas_per_sm : 1));
    if (b.rna)
        detect_tiles_kernel<1><<<grid, NT, sizeof(DetectSmem), st>>>(b, n_tiles, pa_out, sc.bitmap, sc.st_begin,
                                                                      sc.st_end, sc.wit_min, sc.wit_max, seq_flag, sc.tile_read0);
    else
        detect_tiles_kernel<0><<<grid, NT, sizeof(DetectSmem), st>>>(b, n_tiles, pa_out, sc.bitmap, sc.st_begin,
                                                                      sc.st_end, sc.wit_min, sc.wit_max, seq_flag, sc.tile_read0);
    return 1;
}

int launch_verify_tiles(const DevBatch& b, Scratch& sc, uint32_t* seq_flag, uint32_t* fixups, int sm_count,
                        cudaStream_t st) {
    const uint32_t n_tiles = fast_tiles_for(b.span);
    verify_tiles_kernel<<<grid_cap(n_tiles, 256, sm_count * 8), 256, 0, st>>>(b, n_tiles, sc.st_begin, sc.st_end,
                                                                             seq_flag, fixups);
    return 1;
}

int launch_build_seq_list(const DevBatch& b, Scratch& sc, uint32_t* seq_flag, int force_all, int sm_count,
                          cudaStream_t st) {
    build_seq_list_kernel<<<grid_cap(b.n_reads, 256, sm_count * 8), 256, 0, st>>>(
        b, sc.wit_min, sc.wit_max, seq_flag, force_all, sc.seq_list, sc.seq_sbase, sc.seq_count, sc.cursor, sc.gen_cap,
        sc.status, sc.counters);
    return 1;
}

int launch_rank_events(const DevBatch& b, Scratch& sc, uint64_t* ev_off, int sm_count, cudaStream_t st) {
    const uint32_t n_tiles = fast_tiles_for(b.span);
    count_tile_bits_kernel<<<grid_cap((uint64_t)n_tiles * 32, 256, sm_count * 8), 256, 0, st>>>(sc.bitmap, n_tiles,
                                                                                               sc.tile_cnt);
    int n = 1 + launch_scan_u32(sc.tile_cnt, n_tiles, sc.tile_base, reinterpret_cast<uint64_t*>(sc.counters), sc, st);
    read_event_offsets_kernel<<<grid_cap((uint64_t)b.n_reads + 1, 256, sm_count * 8), 256, 0, st>>>(
        b, sc.bitmap, sc.tile_base, n_tiles, ev_off);
    return n + 1;
}

int launch_fast_emit(const DevBatch& b, Scratch& sc, uint64_t ev_cap, uint32_t* ev_start, float* ev_mean,
                     float* ev_stdv, const uint32_t* fixups, int sm_count, cudaStream_t st) {
    const uint32_t n_tiles = fast_tiles_for(b.span);
    const int ctas_per_sm = (int)((227u * 1024u) / (sizeof(EmitSmem) + 1024u));
    emit_tiles_kernel<<<grid_cap(n_tiles, 1, sm_count * ctas_per_sm), ENT, sizeof(EmitSmem), st>>>(
        b, n_tiles, sc.bitmap, sc.tile_base, ev_cap, ev_start, ev_mean, ev_stdv, sc.status, sc.tile_read0);
    sum_fixups_kernel<<<grid_cap(b.n_reads, 256, sm_count * 4), 256, 0, st>>>(b.n_reads, fixups, sc.counters);
    return 2;
}

}  // namespace sgpu
