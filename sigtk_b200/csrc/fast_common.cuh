// fast_common.cuh -- device helpers shared by the kernels of the tiled fast path (emit.cu / detect.cu)
#pragma once
#include "kernels.cuh"

namespace sgpu {
namespace {

constexpr int T = FAST_TILE;            // samples per tile of the event-start bitmap / of emit_tiles_kernel
constexpr int SEG_MAX = 120;            // reads intersecting one macro tile (segment indices are stored as bytes)
constexpr int NONE = INT_MIN;

// Geometry of detect_tiles_kernel. One CTA of DNT threads streams over a MACRO TILE of MS = NCHK*64 samples:
// NPASS passes of PASS = DNT*4 staged samples each produce PC = PASS-2*PH t-statistic positions, so the t arrays
// of one macro tile hold TSPAN = NPASS*PC positions v in [0, TSPAN), v = flat - (macro start - W).
// Chunk c owns v in [W+L*c, W+L*c+L), warms up from v = L*c and may run out to W+L*c+L+R <= TSPAN.
constexpr int DNT = 512;
constexpr int PASS = DNT * 4;           // 4 staged samples per thread and pass
constexpr int PH = 16;                  // halo of a pass on both sides (>= w2 + 2)
constexpr int PC = PASS - 2 * PH;
constexpr int NPASS = 8;
constexpr int TSPAN = NPASS * PC;

template <int RNA>
struct Geo {
    static constexpr int w1 = RNA ? 7 : 3;
    static constexpr int w2 = RNA ? 14 : 6;
    static constexpr int W = RNA ? 128 : 32;   // detector warm-up before a chunk
    static constexpr int R = RNA ? 64 : 32;    // detector run-out after a chunk
    static constexpr int L = RNA ? 64 : 32;    // detector chunk length (one chunk per thread)
    static constexpr int NCHK = (TSPAN - W - R) / L;
    static constexpr int MS = NCHK * L;        // samples per macro tile
    static_assert(NCHK <= DNT && PH >= w2 + 2 && (W % 8) == 0, "geometry");
};

struct Seg {           // one read intersecting the staged region, in region (t-index) coordinates
    int u0;            // region index of the read's first sample (may be negative)
    uint32_t len;
    uint32_t read;
    float off, unit;
};

__device__ __forceinline__ int pad8(int u) { return u + (u >> 3); }     // doubles: 8-sample groups, stride 9
__device__ __forceinline__ int ix4(int u) { return u + (u >> 2); }      // doubles of detect_tiles_kernel: 4-sample groups, stride 5
__device__ __forceinline__ int pad32(int u) { return u + (u >> 5); }    // floats: chunk starts land in distinct banks

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
__device__ __noinline__ int collect_segments(const DevBatch& b, long long lo, long long hi, long long region_start, Seg* segs,
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

// Segmented inclusive prefix sums of x and x*x over the staged region, ITEMS consecutive samples per thread.
// Writes sS/sQ (padded); x outside reads counts as 0; the sums restart at every read start.
template <int NTHREADS, int ITEMS>
__device__ __forceinline__ void region_prefix(const float (&x)[ITEMS], bool starts_read, double* sS, double* sQ,
                                              double* warp_v, double* warp_c, int* warp_f) {
    constexpr int NW = NTHREADS / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double s[ITEMS], q[ITEMS];
    double as = 0.0, aq = 0.0;
#pragma unroll
    for (int m = 0; m < ITEMS; m++) {
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
    if (wid == 0) {  // carry into every warp: segmented exclusive scan of the warp totals by one warp
        double ts = lane < NW ? warp_v[2 * lane] : 0.0, tq = lane < NW ? warp_v[2 * lane + 1] : 0.0;
        int tf = lane < NW ? warp_f[lane] : 0;
#pragma unroll
        for (int o = 1; o < NW; o <<= 1) {
            const double us = __shfl_up_sync(0xffffffffu, ts, o);
            const double uq = __shfl_up_sync(0xffffffffu, tq, o);
            const int uf = __shfl_up_sync(0xffffffffu, tf, o);
            if (lane >= o) {
                if (!tf) { ts = __dadd_rn(ts, us); tq = __dadd_rn(tq, uq); }
                tf |= uf;
            }
        }
        const double xs = __shfl_up_sync(0xffffffffu, ts, 1), xq = __shfl_up_sync(0xffffffffu, tq, 1);
        if (lane < NW) { warp_c[2 * lane] = lane ? xs : 0.0; warp_c[2 * lane + 1] = lane ? xq : 0.0; }
    }
    __syncthreads();
    const double cs = warp_c[2 * wid], cq = warp_c[2 * wid + 1];  // carry from the warps before this one
    double bs, bq;  // exclusive prefix of this thread
    if (starts_read) { bs = 0.0; bq = 0.0; }
    else if (ef) { bs = es; bq = eq; }
    else { bs = __dadd_rn(cs, es); bq = __dadd_rn(cq, eq); }
    const int base = ITEMS == 8 ? pad8(threadIdx.x * 8) : ix4(threadIdx.x * 4);
#pragma unroll
    for (int m = 0; m < ITEMS; m++) {
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
    const unsigned char* grp;   // segment of every 8-sample group, index (v + PH) >> 3; 255 = none
    const Seg* segs;
    int lo, hi;       // region range of the peaks this chunk owns (hi - lo <= 64)

    // run samples [a, b) through both detectors; emissions owned by the chunk set bit (pos - lo) of `mask`
    __device__ __noinline__ void run(DetPair& p, int a, int b, unsigned long long& mask) const {
        using G = Geo<RNA>;
        const DetParams prm = det_params(RNA);
        int sidx = -2, su0 = 0, send = 0;
        for (int u = a; u < b; u++) {
            if (sidx == -2 || (u & 7) == 0) {
                sidx = grp[(u + PH) >> 3];
                if (sidx != 255) { su0 = segs[sidx].u0; send = (int)min((long long)su0 + (long long)segs[sidx].len, (long long)TSPAN); }
            }
            if (sidx == 255 || u >= send) continue;         // alignment gap
            if (u == su0) { det_set(p.s, u); det_set(p.l, u); }  // first sample of a read: initial state (516-536)
            const float c1 = t1[pad32(u)], c2 = t2[pad32(u)];
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
            const int sidx = grp[(u + PH) >> 3];
            if (sidx == 255) return true;                           // the read ended: pending peaks are dropped
            const int su0 = segs[sidx].u0;
            if (u == su0 || (long long)u - su0 >= (long long)segs[sidx].len) return true;
            run(p, u, u + 1, mask);
            u++;
        }
        return true;
    }
};

// One sample of one detector without branches (same transitions as det_step, events.c:387-437), for a sample
// that lies inside a read and is not its first sample. Returns true when a peak is emitted; *pos is its position.
template <bool SHORT, int RNA>
__device__ __forceinline__ bool step_one(Det& d, Det& lng, int u, float c, int* pos) {
    using G = Geo<RNA>;
    constexpr float thr = SHORT ? (RNA ? 2.5f : 1.4f) : 9.0f, h = RNA ? 1.0f : 0.2f;
    constexpr int w = SHORT ? G::w1 : G::w2;
    const bool act = d.mt < u, none = d.pp == NONE;
    const bool lt = c < d.pv, gt = c > d.pv;
    const bool rise = __fsub_rn(c, d.pv) > h;
    const float pv2 = gt ? c : d.pv;
    const int pp2 = gt ? u : d.pp;
    const bool big = pv2 > thr;
    const bool valid2 = (d.valid != 0) | (big & (__fsub_rn(pv2, c) > h));
    const bool emit = valid2 & ((int)((unsigned)u - (unsigned)pp2) > w / 2);
    const bool in1 = act & none, in2 = act & !none;
    if (SHORT) {  // the short detector dominates the long one (414-422)
        const bool maskl = in2 & big;
        lng.mt = maskl ? pp2 + G::w1 : lng.mt;
        lng.pp = maskl ? NONE : lng.pp;
        lng.pv = maskl ? FLT_MAX : lng.pv;
        lng.valid = maskl ? 0 : lng.valid;
    }
    const bool e = in2 & emit;
    const bool set_c = (in1 & (lt | rise)) | (in2 & (gt | emit));
    const bool set_u = (in1 & !lt & rise) | (in2 & gt & !emit);
    *pos = pp2;
    d.pv = set_c ? c : d.pv;
    d.pp = e ? NONE : (set_u ? u : d.pp);
    d.valid = in2 ? ((valid2 & !emit) ? 1 : 0) : d.valid;
    return e;
}

}  // namespace
}  // namespace sgpu
