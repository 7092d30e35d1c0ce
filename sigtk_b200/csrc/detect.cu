// detect.cu -- detect_tiles_kernel: the first (and dominant) kernel of the fast path.
//
// One persistent CTA per SM streams over MACRO TILES of MS ~ 16k samples of the flat int16 array:
//
//   for each of NPASS passes (2048 staged samples, TMA bulk copy, double buffered, mbarrier completion):
//     B  pA conversion fused into the load (misc.c:26-29; pA optionally stored, 128-bit stores) and segmented
//        FP64 inclusive prefix sums of x and x*x (events.c:293-303) over the pass,
//     C  both window t-statistics (events.c:315-364) for the pass's 2016 core positions -> shared t arrays
//   D  the dual peak detector (events.c:371-443), one 64-sample chunk per THREAD (all warps busy): every chunk
//      starts W samples early from a cold state; the state at the chunk start is compared with the previous
//      chunk's end state and mismatching chunks are re-run from the true state; peaks are owned by position, so a
//      chunk runs on past its end until its pending peaks are resolved.
//   out: event-start bits of the macro tile, the detector state at both ends of the macro tile (verified across
//        macro tiles by verify_tiles_kernel), the exact-sum witness per read.
#include "fast_common.cuh"

namespace sgpu {

namespace {

struct DetectSmem {
    alignas(16) int16_t raw[2][PASS];
    alignas(8) uint64_t bar[2];
    double sS[PASS + PASS / 4 + 8];          // the detector phase reuses this array for the chunk end states
    double sQ[PASS + PASS / 4 + 8];
    float t1[TSPAN + TSPAN / 32 + 4];
    float t2[TSPAN + TSPAN / 32 + 4];
    Seg segbuf[2][SEG_MAX];                  // segments of this macro tile / prefetched for the next one
    uint32_t wmin[SEG_MAX], wmax[SEG_MAX];
    double warp_v[2 * (DNT / 32)];
    double warp_c[2 * (DNT / 32)];           // carry into every warp (region_prefix)
    int warp_f[DNT / 32];
    uint32_t kind[TSPAN / 16];               // 2 bits per position: 0 inside a read, 1 first sample of a read, 2 gap
    uint32_t bits[2 * DNT];                  // event-start bits of the macro tile
    unsigned char flag[TSPAN];               // 1 = a peak was emitted at this position (by the chunk that owns it)
    unsigned char grp[(TSPAN + 2 * PH) / 8]; // segment of every 8-sample group, index (v + PH) >> 3; 255 = none
    uint32_t pass_min;                       // min nonzero |pA| bit pattern of the staged pass
    int nsegbuf[2], ovfbuf[2];
};
static_assert(sizeof(double) * (PASS + PASS / 4 + 8) >= sizeof(int) * 8 * DNT, "end states alias sS");

}  // namespace

template <int RNA>
__global__ void __launch_bounds__(DNT, 1) detect_tiles_kernel(DevBatch b, uint32_t n_macro, uint64_t bitmap_words,
                                                             float* __restrict__ pa_out,
                                                             uint32_t* __restrict__ bitmap, int* __restrict__ st_begin,
                                                             int* __restrict__ st_end, uint32_t* __restrict__ wit_min,
                                                             uint32_t* __restrict__ wit_max,
                                                             uint32_t* __restrict__ seq_flag,
                                                             const uint32_t* __restrict__ macro_read0) {
    using G = Geo<RNA>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DetectSmem& sm = *reinterpret_cast<DetectSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const long long span = (long long)b.span;

    // flat position of staged sample 0 of pass `ps` of macro tile `m`
    auto pass_start = [&](uint32_t m, int ps) { return (long long)m * G::MS - G::W - PH + (long long)ps * PC; };
    auto issue_load = [&](uint32_t m, int ps, int buf) {  // one thread: bulk copy of the valid part of the pass
        const long long rs = pass_start(m, ps);
        const long long lo = rs < 0 ? 0 : rs;
        long long hi = rs + PASS;
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
        if (blockIdx.x < n_macro) issue_load(blockIdx.x, 0, 0);
    }
    __syncthreads();

    uint32_t n_loads = 0;  // bulk copies consumed so far by this CTA (buffer = n_loads & 1)
    int sb = 0;            // which segment buffer describes the current macro tile
    auto fetch_segments = [&](uint32_t mm, int which) {  // one thread
        const long long f0 = (long long)mm * G::MS - G::W;
        int ovf = 0;
        sm.nsegbuf[which] = collect_segments(b, f0 - PH, f0 + TSPAN + PH, f0, sm.segbuf[which], &ovf, seq_flag,
                                             macro_read0[mm]);
        sm.ovfbuf[which] = ovf;
    };
    if (tid == DNT - 1 && blockIdx.x < n_macro) fetch_segments(blockIdx.x, 0);
    for (uint32_t m = blockIdx.x; m < n_macro; m += gridDim.x, sb ^= 1) {
        const long long v0_flat = (long long)m * G::MS - G::W;  // flat position of t index v = 0
        Seg* const segs = sm.segbuf[sb];
        for (int k = tid; k < SEG_MAX; k += DNT) { sm.wmin[k] = 0xffffffffu; sm.wmax[k] = 0u; }
        for (int k = tid; k < TSPAN / 4; k += DNT) reinterpret_cast<uint32_t*>(sm.flag)[k] = 0u;
        __syncthreads();  // the segments were fetched before the previous macro tile's last barrier (or just above)
        const int nseg = sm.nsegbuf[sb];
        for (int g = tid; g < (TSPAN + 2 * PH) / 8; g += DNT) {
            const int s = group_segment(segs, nseg, g * 8 - PH);
            sm.grp[g] = (unsigned char)(s < 0 ? 255 : s);
        }
        __syncthreads();
        for (int j = tid; j < TSPAN / 16; j += DNT) {  // position kinds, 16 positions per word
            uint32_t kw = 0u;
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const int vg = j * 16 + half * 8;
                const int sidx = sm.grp[(vg + PH) >> 3];
                uint32_t k8 = 0xaaaau;  // eight gaps
                if (sidx != 255) {
                    const int done = vg - segs[sidx].u0;  // >= 0
                    const uint32_t len = segs[sidx].len;
                    k8 = 0u;
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const uint32_t i = (uint32_t)(done + q);
                        k8 |= (i == 0u ? 1u : (i < len ? 0u : 2u)) << (2 * q);
                    }
                }
                kw |= k8 << (16 * half);
            }
            sm.kind[j] = kw;
        }
        // (the barrier after the first pass's phase B orders these writes before the detector phase)

        // ---- passes: B (pA + prefix sums) and C (t-statistics) ---------------------------------------------------------
        for (int ps = 0; ps < NPASS; ps++, n_loads++) {
            const int buf = n_loads & 1;
            if (tid == 0) {  // prefetch the next pass (its buffer was consumed before the last barrier)
                fence_proxy_async();
                if (ps + 1 < NPASS) issue_load(m, ps + 1, buf ^ 1);
                else if (m + gridDim.x < n_macro) issue_load(m + gridDim.x, 0, buf ^ 1);
            }
            mbar_wait(&sm.bar[buf], (n_loads >> 1) & 1);
            const int vb = ps * PC - PH;  // t index of staged sample 0 of this pass
            // interior pass: every staged sample lies inside one read and none is its first sample
            bool interior = false;
            {
                const int s0 = sm.grp[(vb + PH) >> 3];
                if (s0 != 255) interior = segs[s0].u0 < vb && (long long)segs[s0].u0 + (long long)segs[s0].len >= vb + PASS;
            }
            if (tid == 0) sm.pass_min = 0xffffffffu;
            __syncthreads();
            {
                const int r0 = tid * 4;
                const int vg = vb + r0;
                int sidx = sm.grp[(vg + PH) >> 3];
                float x[4];
                bool starts = false;
                uint32_t mn = 0xffffffffu, mx = 0u, own_mn = 0xffffffffu;
                Seg sg;
                if (sidx != 255) sg = segs[sidx];
                const bool ovf = sm.ovfbuf[sb] != 0;
                if (ovf) {
                    // more reads than the segment list holds: all of them go to the sequential-order kernels, but
                    // their pA is still stored here, so every thread looks its read up in global memory
                    sidx = 255;
                    const long long fp = v0_flat + vg;
                    if (fp >= 0 && fp < span) {
                        const uint32_t r = find_read(b.read_off, b.n_reads, (uint64_t)fp);
                        const long long rs = (long long)b.read_off[r];
                        if (fp - rs < (long long)b.read_len[r]) {
                            sg.u0 = (int)(rs - v0_flat); sg.len = b.read_len[r]; sg.read = r;
                            sg.off = b.offset[r]; sg.unit = b.unit[r];
                            sidx = 254;
                        }
                    }
                }
                if (sidx != 255) {
                    starts = (sg.u0 == vg);
                    const int2 rawv = *reinterpret_cast<const int2*>(&sm.raw[buf][r0]);
                    const int v[4] = {(int)(int16_t)(rawv.x & 0xffff), rawv.x >> 16, (int)(int16_t)(rawv.y & 0xffff),
                                      rawv.y >> 16};
                    const uint32_t done = (uint32_t)(vg - sg.u0);
                    const uint32_t left = done < sg.len ? sg.len - done : 0u;  // samples of the read from vg on
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const float xv = __fmul_rn(__fadd_rn((float)v[k], sg.off), sg.unit);
                        x[k] = ((uint32_t)k < left) ? xv : 0.0f;
                        const uint32_t a = __float_as_uint(x[k]) & 0x7fffffffu;
                        mx = max(mx, a);
                        mn = min(mn, a ? a : 0xffffffffu);
                    }
                    // every sample is owned by exactly one (macro tile, pass): witness + pA store happen there
                    const bool own = r0 >= PH && r0 < PH + PC && vg >= G::W && vg < G::W + G::MS;
                    if (own && pa_out)
                        *reinterpret_cast<float4*>(pa_out + (v0_flat + vg)) = make_float4(x[0], x[1], x[2], x[3]);
                    if (!own || ovf) { mx = 0u; }
                    own_mn = (own && !ovf) ? mn : 0xffffffffu;
                } else {
#pragma unroll
                    for (int k = 0; k < 4; k++) x[k] = 0.0f;
                }
                {   // exact-sum witness of the owned samples and the pass minimum: one shared atomic per warp
                    const int s0 = __shfl_sync(0xffffffffu, sidx, 0);
                    const bool uni = __all_sync(0xffffffffu, sidx == s0) && !ovf;
                    const uint32_t pmn = __reduce_min_sync(0xffffffffu, mn);
                    if (uni) {
                        const uint32_t wmn = __reduce_min_sync(0xffffffffu, own_mn);
                        const uint32_t wmx = __reduce_max_sync(0xffffffffu, mx);
                        if ((tid & 31) == 0 && s0 != 255 && (wmx | (uint32_t)(wmn != 0xffffffffu))) {
                            atomicMin(&sm.wmin[s0], wmn);
                            atomicMax(&sm.wmax[s0], wmx);
                        }
                    } else if (sidx != 255 && (mx | (uint32_t)(own_mn != 0xffffffffu))) {
                        atomicMin(&sm.wmin[sidx], own_mn);
                        atomicMax(&sm.wmax[sidx], mx);
                    }
                    if ((tid & 31) == 0) atomicMin(&sm.pass_min, pmn);
                }
                region_prefix<DNT, 4>(x, starts, sm.sS, sm.sQ, sm.warp_v, sm.warp_c, sm.warp_f);
            }
            __syncthreads();
            if (interior && sm.pass_min >= 0x21800000u) {  // all staged |pA| >= 2^-60: lean loop, no per-position checks
#pragma unroll 2
                for (int r = PH + tid; r < PH + PC; r += DNT) {
                    const double s_i = sm.sS[ix4(r - 1)], q_i = sm.sQ[ix4(r - 1)];
                    const float r1 = tstat_fast<G::w1, true>(
                        __dsub_rn(s_i, sm.sS[ix4(r - G::w1 - 1)]), __dsub_rn(q_i, sm.sQ[ix4(r - G::w1 - 1)]),
                        __dsub_rn(sm.sS[ix4(r + G::w1 - 1)], s_i), __dsub_rn(sm.sQ[ix4(r + G::w1 - 1)], q_i));
                    const float r2 = tstat_fast<G::w2, true>(
                        __dsub_rn(s_i, sm.sS[ix4(r - G::w2 - 1)]), __dsub_rn(q_i, sm.sQ[ix4(r - G::w2 - 1)]),
                        __dsub_rn(sm.sS[ix4(r + G::w2 - 1)], s_i), __dsub_rn(sm.sQ[ix4(r + G::w2 - 1)], q_i));
                    sm.t1[pad32(vb + r)] = r1;
                    sm.t2[pad32(vb + r)] = r2;
                }
            } else {
                for (int r = PH + tid; r < PH + PC; r += DNT) {
                    const int v = vb + r;
                    float r1 = 0.0f, r2 = 0.0f;
                    const int sidx = sm.grp[(v + PH) >> 3];
                    if (sidx != 255) {
                        const int su0 = segs[sidx].u0;
                        const uint32_t n = segs[sidx].len;
                        const uint32_t i = (uint32_t)(v - su0);
                        if (i < n) {
                            const double s_i = sm.sS[ix4(r - 1)], q_i = sm.sQ[ix4(r - 1)];  // i >= w >= 2 where used
                            if (n >= 2u * G::w1 && i >= (uint32_t)G::w1 && i + G::w1 <= n) {
                                const bool first = (i == (uint32_t)G::w1);
                                const double sl = first ? 0.0 : sm.sS[ix4(r - G::w1 - 1)];
                                const double ql = first ? 0.0 : sm.sQ[ix4(r - G::w1 - 1)];
                                r1 = tstat_fast<G::w1>(__dsub_rn(s_i, sl), __dsub_rn(q_i, ql),
                                                       __dsub_rn(sm.sS[ix4(r + G::w1 - 1)], s_i),
                                                       __dsub_rn(sm.sQ[ix4(r + G::w1 - 1)], q_i));
                            }
                            if (n >= 2u * G::w2 && i >= (uint32_t)G::w2 && i + G::w2 <= n) {
                                const bool first = (i == (uint32_t)G::w2);
                                const double sl = first ? 0.0 : sm.sS[ix4(r - G::w2 - 1)];
                                const double ql = first ? 0.0 : sm.sQ[ix4(r - G::w2 - 1)];
                                r2 = tstat_fast<G::w2>(__dsub_rn(s_i, sl), __dsub_rn(q_i, ql),
                                                       __dsub_rn(sm.sS[ix4(r + G::w2 - 1)], s_i),
                                                       __dsub_rn(sm.sQ[ix4(r + G::w2 - 1)], q_i));
                            }
                        }
                    }
                    sm.t1[pad32(v)] = r1;
                    sm.t2[pad32(v)] = r2;
                }
            }
            __syncthreads();
        }
        // witness: one global atomic pair per read of the macro tile
        for (int k = tid; k < nseg; k += DNT) {
            if (sm.wmax[k] | (uint32_t)(sm.wmin[k] != 0xffffffffu)) {
                atomicMin(&wit_min[segs[k].read], sm.wmin[k]);
                atomicMax(&wit_max[segs[k].read], sm.wmax[k]);
            }
        }

        // ---- phase D: the peak detector, one 64-sample chunk per thread ------------------------------------------------
        const bool has_chunk = tid < G::NCHK;
        constexpr int L = G::L;
        const int cs = G::W + tid * L, ce = cs + L;
        const int wa = cs - G::W, wz = ce + G::R;  // everything this thread may look at
        Walker<RNA> wk{sm.t1, sm.t2, sm.grp, segs, cs, ce};
        const DetParams prm = det_params(RNA);
        DetPair p;
        det_set(p.s, wa - 1);  // cold start: wa is the first sample processed
        det_set(p.l, wa - 1);
        unsigned long long mask = 0ull;  // peaks found by the general walker (only used by re-runs)
        Canon begin, end;
        bool ok = true;
        if (tid == DNT - 1 && m + gridDim.x < n_macro) fetch_segments(m + gridDim.x, sb ^ 1);  // idle in this phase
        if (has_chunk) {
            int pos;
            // 16 positions per kind word; wa, cs and ce are multiples of 32
            auto walk16 = [&](int u0, bool record) {
                const uint32_t kw = sm.kind[u0 >> 4];
                const float* const a1 = sm.t1 + pad32(u0);  // u0 is a multiple of 16: the 16 positions are contiguous
                const float* const a2 = sm.t2 + pad32(u0);
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    const int u = u0 + q;
                    const uint32_t k = (kw >> (2 * q)) & 3u;
                    if (k) {                       // rare: alignment gap (skip) or first sample of a read (reset)
                        if (k == 2u) continue;
                        det_set(p.s, u);
                        det_set(p.l, u);
                    }
                    const bool e1 = step_one<true, RNA>(p.s, p.l, u, a1[q], &pos);
                    if (record && e1 && pos >= cs) sm.flag[pos] = 1;
                    const bool e2 = step_one<false, RNA>(p.l, p.l, u, a2[q], &pos);
                    if (record && e2 && pos >= cs) sm.flag[pos] = 1;
                }
            };
            for (int u0 = wa; u0 < cs; u0 += 16) walk16(u0, false);
            begin = canon(p, cs);
            for (int u0 = cs; u0 < ce; u0 += 16) walk16(u0, true);
            end = canon(p, ce);
            // run on until the pending peaks this chunk owns are resolved (or the read ends)
            DetPair r = p;
            int u = ce;
            while (wk.pending(r.s, prm.thr1) || wk.pending(r.l, prm.thr2)) {
                if (u >= wz) { ok = false; break; }
                if ((sm.kind[u >> 4] >> (2 * (u & 15))) & 3u) break;  // the read ended: pending peaks are dropped
                if (step_one<true, RNA>(r.s, r.l, u, sm.t1[pad32(u)], &pos) && pos >= cs && pos < ce) sm.flag[pos] = 1;
                if (step_one<false, RNA>(r.l, r.l, u, sm.t2[pad32(u)], &pos) && pos >= cs && pos < ce) sm.flag[pos] = 1;
                u++;
            }
        }
        const Canon macro_begin = begin;  // chunk 0: speculative state at the macro tile start
        int (*endst)[8] = reinterpret_cast<int (*)[8]>(sm.sS);  // the prefix sums are dead during this phase
        // compare with the previous chunk's end state; re-run mismatching chunks from the true state
        for (int round = 0; round < G::NCHK; round++) {
            if (has_chunk) {
#pragma unroll
                for (int k = 0; k < 8; k++) endst[tid][k] = end.v[k];
            }
            __syncthreads();
            bool mism = false;
            Canon prev;
            if (has_chunk && tid > 0) {
#pragma unroll
                for (int k = 0; k < 8; k++) prev.v[k] = endst[tid - 1][k];
                mism = !canon_eq(prev, begin);
            }
            if (!__syncthreads_or(mism ? 1 : 0)) break;
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
                for (int k = 0; k < L / 4; k++) reinterpret_cast<uint32_t*>(sm.flag + cs)[k] = 0u;  // cs % 4 == 0
                wk.run(p, cs, ce, mask);
                end = canon(p, ce);
                ok = wk.run_out(p, ce, mask);
            }
        }
        if (has_chunk) {
            // the chunk's L flag bytes -> bitmap word(s), merged with the slow walker's mask
            unsigned long long bitsv = mask;
            const uint32_t* fw = reinterpret_cast<const uint32_t*>(sm.flag + cs);
#pragma unroll
            for (int k = 0; k < L / 4; k++) {
                const uint32_t w4 = fw[k];  // four flag bytes (0/1) -> four bits
                const uint32_t nib = (w4 & 1u) | ((w4 >> 7) & 2u) | ((w4 >> 14) & 4u) | ((w4 >> 21) & 8u);
                bitsv |= (unsigned long long)nib << (4 * k);
            }
            if (L == 64) {
                sm.bits[2 * tid] = (uint32_t)bitsv;
                sm.bits[2 * tid + 1] = (uint32_t)(bitsv >> 32);
            } else {
                sm.bits[tid] = (uint32_t)bitsv;
            }
            if (!ok) {  // run-out cap hit: let the sequential-order kernels do the read that contains ce-1
                const int sidx = sm.grp[(ce - 1 + PH) >> 3];
                if (sidx != 255) seq_flag[segs[sidx].read] = 1u;
            }
            if (tid == 0) {
#pragma unroll
                for (int k = 0; k < 8; k++) st_begin[(size_t)m * 8 + k] = macro_begin.v[k];
            }
            if (tid == G::NCHK - 1) {
#pragma unroll
                for (int k = 0; k < 8; k++) st_end[(size_t)m * 8 + k] = end.v[k];
            }
        }
        __syncthreads();
        // event 0 of every read starts at its first sample (events.c:490-497)
        for (int k = tid; k < nseg; k += DNT) {
            const int c = segs[k].u0 - G::W;
            if (c >= 0 && c < G::MS) atomicOr(&sm.bits[c >> 5], 1u << (c & 31));
            if (sm.ovfbuf[sb]) seq_flag[segs[k].read] = 1u;
        }
        __syncthreads();
        for (int k = tid; k < G::MS / 32; k += DNT) {
            const uint64_t wi = (uint64_t)m * (G::MS / 32) + k;
            if (wi < bitmap_words) bitmap[wi] = sm.bits[k];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Macro tile m (m >= 1) started its first chunk from a speculative state; it must equal the end state of macro
// tile m-1 whenever the boundary lies strictly inside a read.
__global__ void __launch_bounds__(256) verify_tiles_kernel(DevBatch b, uint32_t n_macro, uint32_t macro_samples,
                                                           const int* __restrict__ st_begin,
                                                           const int* __restrict__ st_end,
                                                           uint32_t* __restrict__ seq_flag,
                                                           uint32_t* __restrict__ fixups) {
    for (uint32_t t = 1 + blockIdx.x * blockDim.x + threadIdx.x; t < n_macro; t += gridDim.x * blockDim.x) {
        const uint64_t p = (uint64_t)t * macro_samples;
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

// first read that can intersect the staged range of every macro tile
__global__ void __launch_bounds__(256) macro_read0_kernel(DevBatch b, uint32_t n_macro, uint32_t macro_samples,
                                                          uint32_t* __restrict__ macro_read0) {
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_macro; t += gridDim.x * blockDim.x) {
        const long long p = (long long)t * macro_samples - 160;  // W + PH <= 144
        macro_read0[t] = find_read(b.read_off, b.n_reads, (uint64_t)(p < 0 ? 0 : p));
    }
}

static inline int grid_cap(uint64_t work, int block, int max_blocks) {
    uint64_t g = (work + block - 1) / block;
    if (g < 1) g = 1;
    if (g > (uint64_t)max_blocks) g = max_blocks;
    return (int)g;
}

static uint32_t macro_samples(int rna) { return rna ? (uint32_t)Geo<1>::MS : (uint32_t)Geo<0>::MS; }
// macro tiles must cover every word of the 2048-tiled bitmap that emit_tiles_kernel reads
static uint32_t macro_tiles_for(uint64_t span, int rna) {
    const uint64_t covered = (uint64_t)fast_tiles_for(span) * FAST_TILE;
    const uint32_t ms = macro_samples(rna);
    return (uint32_t)((covered + ms - 1) / ms);
}

int detect_configure() {
    cudaError_t e = cudaFuncSetAttribute(detect_tiles_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(DetectSmem));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(detect_tiles_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(DetectSmem));
    return e == cudaSuccess ? 0 : -1;
}

int launch_fast_detect(const DevBatch& b, Scratch& sc, float* pa_out, uint32_t* seq_flag, uint32_t* fixups,
                       int sm_count, cudaStream_t st) {
    const uint32_t n_macro = macro_tiles_for(b.span, b.rna);
    const uint64_t words = (uint64_t)fast_tiles_for(b.span) * (FAST_TILE / 32);
    macro_read0_kernel<<<grid_cap(n_macro, 256, sm_count * 4), 256, 0, st>>>(b, n_macro, macro_samples(b.rna),
                                                                            sc.macro_read0);
    const int grid = grid_cap(n_macro, 1, sm_count);
    if (b.rna)
        detect_tiles_kernel<1><<<grid, DNT, sizeof(DetectSmem), st>>>(b, n_macro, words, pa_out, sc.bitmap, sc.st_begin,
                                                                       sc.st_end, sc.wit_min, sc.wit_max, seq_flag,
                                                                       sc.macro_read0);
    else
        detect_tiles_kernel<0><<<grid, DNT, sizeof(DetectSmem), st>>>(b, n_macro, words, pa_out, sc.bitmap, sc.st_begin,
                                                                       sc.st_end, sc.wit_min, sc.wit_max, seq_flag,
                                                                       sc.macro_read0);
    return 2;
}

int launch_verify_tiles(const DevBatch& b, Scratch& sc, uint32_t* seq_flag, uint32_t* fixups, int sm_count,
                        cudaStream_t st) {
    const uint32_t n_macro = macro_tiles_for(b.span, b.rna);
    verify_tiles_kernel<<<grid_cap(n_macro, 256, sm_count * 4), 256, 0, st>>>(b, n_macro, macro_samples(b.rna),
                                                                             sc.st_begin, sc.st_end, seq_flag, fixups);
    return 1;
}

}  // namespace sgpu
