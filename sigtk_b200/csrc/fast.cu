// fast.cu -- the tiled fast path of the sigtk B200 hot path (sm_100a); detect_tiles_kernel lives in detect.cu.
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
#include "fast_common.cuh"

namespace sgpu {

// ---------------------------------------------------------------------------------------------------------------
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
        // the walker's unguarded shortcuts need unit > 0 (no -0 sums) and 32-bit positions with headroom
        const float unit = b.unit[r];
        if (!(unit > 0.0f && unit <= FLT_MAX) || n >= (1u << 30)) flag = true;
        if (!flag && n > 0) {
            const uint32_t mn = wit_min[r], mx = wit_max[r];
            // every sample must be a positive float >= 2^-60 (the range of the walker's shortcuts) ...
            if (mn < 0x21800000u || mn > mx) flag = true;
            else {  // ... and the sums of x and x*x must be exact whatever the order
                const uint32_t log2n = n > 1 ? 32u - __clz(n - 1) : 0u;
                const float fmn = __uint_as_float(mn), fmx = __uint_as_float(mx);
                const uint32_t qmn = __float_as_uint(__fmul_rn(fmn, fmn)), qmx = __float_as_uint(__fmul_rn(fmx, fmx));
                if (!(sums_exact(mn >> 23, mx >> 23, log2n) && sums_exact(qmn >> 23, qmx >> 23, log2n))) flag = true;
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
    Seg segs[SEG_MAX];       // reads intersecting the EMT tiles of this round, positions relative to the first tile
    Seg gseg[T / 8];         // the read owning every 8-sample group of the current tile (len 0 = none), tile coords
    double warp_v[2 * (ENT / 32)];
    double warp_c[2 * (ENT / 32)];
    int warp_f[ENT / 32];
    uint32_t bits[T / 32];
    uint32_t excl[T / 32];
    int nseg, overflow;
    int spill_u;            // tile index of the start of the event that runs past the tile end, or -1
    unsigned long long spill_k;
    // the open event carried from tile to tile inside one group of EMT tiles
    int carry_on;
    unsigned long long carry_k;
    double carry_s, carry_q;
    long long carry_start, carry_read_end;  // flat positions
    float carry_off, carry_unit;
};

__global__ void __launch_bounds__(ENT) emit_tiles_kernel(DevBatch b, uint32_t n_tiles, const uint32_t* __restrict__ bitmap,
                                                         const uint64_t* __restrict__ tile_base, uint64_t ev_cap,
                                                         uint32_t* __restrict__ ev_start, float* __restrict__ ev_mean,
                                                         float* __restrict__ ev_stdv, int* __restrict__ status,
                                                         const uint32_t* __restrict__ tile_read0) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EmitSmem& sm = *reinterpret_cast<EmitSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    constexpr int EMT = 16;  // consecutive tiles per segment collection
    for (uint32_t tile0 = blockIdx.x * EMT; tile0 < n_tiles; tile0 += gridDim.x * EMT) {
      __syncthreads();
      if (tid == 0) {
          int ovf = 0;
          const long long f0 = (long long)tile0 * T;
          sm.nseg = collect_segments(b, f0, f0 + (long long)EMT * T, f0, sm.segs, &ovf, nullptr, tile_read0[tile0]);
          sm.overflow = ovf;  // more reads than the list holds: every thread looks its read up in global memory
          sm.carry_on = 0;
      }
      const uint32_t tile_end = min(tile0 + EMT, n_tiles);
      // software prefetch: the global loads of tile t+1 are issued before tile t is processed
      auto load_raw = [&](uint32_t t) {
          const long long p = (long long)t * T + threadIdx.x * 8;
          return (t < tile_end && p < (long long)b.span) ? __ldg(reinterpret_cast<const int4*>(b.samples + p))
                                                         : make_int4(0, 0, 0, 0);
      };
      auto load_bits = [&](uint32_t t) { return (t < tile_end && tid < T / 32) ? bitmap[(size_t)t * (T / 32) + tid] : 0u; };
      auto load_base = [&](uint32_t t) { return t < tile_end ? tile_base[t] : 0ull; };
      int4 raw_nx = load_raw(tile0);
      uint32_t bits_nx = load_bits(tile0);
      uint64_t base_nx = load_base(tile0);
      __syncthreads();
      for (uint32_t tile = tile0; tile < tile_end; tile++) {
        const long long ts = (long long)tile * T;
        const int toff = (int)(tile - tile0) * T;
        const int4 rawv = raw_nx;
        const uint64_t base = base_nx;
        if (tid == 0) sm.spill_u = -1;
        if (tid < T / 32) sm.bits[tid] = bits_nx;
        raw_nx = load_raw(tile + 1);
        bits_nx = load_bits(tile + 1);
        base_nx = load_base(tile + 1);
        __syncthreads();
        const int nseg = sm.nseg;
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
        Seg sg;  // the read owning this thread's 8 samples, in tile coordinates (len 0: none)
        sg.len = 0u; sg.u0 = 0; sg.read = 0u; sg.off = 0.0f; sg.unit = 0.0f;
        if (!sm.overflow) {
            const int sidx = group_segment(sm.segs, nseg, toff + u0);
            if (sidx >= 0) { sg = sm.segs[sidx]; sg.u0 -= toff; }
        } else if (ts + u0 < (long long)b.span) {
            const uint32_t r = find_read(b.read_off, b.n_reads, (uint64_t)(ts + u0));
            const long long rs = (long long)b.read_off[r];
            const uint32_t len = b.read_len[r];
            if (ts + u0 - rs < (long long)len) {
                sg.u0 = (int)(rs - ts); sg.len = len; sg.read = r; sg.off = b.offset[r]; sg.unit = b.unit[r];
            }
        }
        sm.gseg[tid] = sg;
        const int sidx = sg.len ? 0 : -1;
        {
            float x[8];
            bool starts = false;
            if (sidx >= 0 && ts + u0 < (long long)b.span) {
                starts = (sg.u0 == u0);
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
            region_prefix<ENT, 8>(x, starts, sm.sS, sm.sQ, sm.warp_v, sm.warp_c, sm.warp_f);
        }
        __syncthreads();
        // every event start bit in this thread's 8 samples
        if (sidx >= 0) {
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
        // events that cross tile ends: carried from tile to tile inside the group by warp 0; only at the end of
        // the group the rest of the event is summed from global memory
        if (tid < 32) {
            if (sm.carry_on) {  // (a) the event carried in: it ends at the first start bit of this tile or at its read's end
                const uint32_t w0 = sm.bits[lane], w1 = sm.bits[lane + 32];
                const uint32_t any0 = __ballot_sync(0xffffffffu, w0 != 0u), any1 = __ballot_sync(0xffffffffu, w1 != 0u);
                long long end = (long long)T + 1;
                if (any0) {
                    const int f = __ffs(any0) - 1;
                    end = (f << 5) + __ffs(__shfl_sync(0xffffffffu, w0, f)) - 1;
                } else if (any1) {
                    const int f = __ffs(any1) - 1;
                    end = ((f + 32) << 5) + __ffs(__shfl_sync(0xffffffffu, w1, f)) - 1;
                }
                const long long read_end = sm.carry_read_end - ts;
                if (end > read_end) end = read_end;
                if (lane == 0) {
                    if (end <= T) {
                        const double as = end > 0 ? sm.sS[pad8((int)end - 1)] : 0.0;
                        const double aq = end > 0 ? sm.sQ[pad8((int)end - 1)] : 0.0;
                        float mean, stdv;
                        event_stats(__dadd_rn(sm.carry_s, as), __dadd_rn(sm.carry_q, aq),
                                    (uint32_t)(ts + end - sm.carry_start), &mean, &stdv);
                        ev_mean[sm.carry_k] = mean;
                        ev_stdv[sm.carry_k] = stdv;
                        sm.carry_on = 0;
                    } else {  // the read covers the whole tile and no event starts in it
                        sm.carry_s = __dadd_rn(sm.carry_s, sm.sS[pad8(T - 1)]);
                        sm.carry_q = __dadd_rn(sm.carry_q, sm.sQ[pad8(T - 1)]);
                    }
                }
                __syncwarp();
            }
            if (sm.spill_u >= 0 && lane == 0) {  // (b) the last event of this tile runs past the tile end
                const int u = sm.spill_u;
                const Seg sg = sm.gseg[u >> 3];
                const bool at0 = (u == sg.u0) || (u == 0);
                sm.carry_s = __dsub_rn(sm.sS[pad8(T - 1)], at0 ? 0.0 : sm.sS[pad8(u - 1)]);
                sm.carry_q = __dsub_rn(sm.sQ[pad8(T - 1)], at0 ? 0.0 : sm.sQ[pad8(u - 1)]);
                sm.carry_k = sm.spill_k;
                sm.carry_start = ts + u;
                sm.carry_read_end = ts + sg.u0 + (long long)sg.len;
                sm.carry_off = sg.off;
                sm.carry_unit = sg.unit;
                sm.carry_on = 1;
            }
            __syncwarp();
            if (sm.carry_on && tile + 1 == tile_end) {  // (c) end of the group: finish the event from global memory
                const long long read_end = sm.carry_read_end;
                const float off = sm.carry_off, unit = sm.carry_unit;
                // the next event start at or after the end of this tile (bounded by the end of the read)
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
                    const int4 rv = __ldg(reinterpret_cast<const int4*>(b.samples + p));
                    int v[8];
                    unpack8(rv, v);
#pragma unroll
                    for (int m = 0; m < 8; m++) {
                        if (p + m < end) {
                            const float xv = __fmul_rn(__fadd_rn((float)v[m], off), unit);
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
                    event_stats(__dadd_rn(sm.carry_s, as), __dadd_rn(sm.carry_q, aq), (uint32_t)(end - sm.carry_start),
                                &mean, &stdv);
                    ev_mean[sm.carry_k] = mean;
                    ev_stdv[sm.carry_k] = stdv;
                    sm.carry_on = 0;
                }
            }
        }
        __syncthreads();
      }
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
    cudaError_t e =
        cudaFuncSetAttribute(emit_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EmitSmem));
    return e == cudaSuccess ? 0 : -1;
}

int launch_init_reads(const DevBatch& b, Scratch& sc, uint32_t* seq_flag, uint32_t* fixups, int sm_count,
                      cudaStream_t st) {
    const uint32_t n_tiles = fast_tiles_for(b.span);
    init_reads_kernel<<<grid_cap(max(b.n_reads, n_tiles), 256, sm_count * 8), 256, 0, st>>>(
        b, n_tiles, sc.wit_min, sc.wit_max, seq_flag, fixups, sc.seq_count, sc.cursor, sc.tile_read0);
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
    emit_tiles_kernel<<<grid_cap((n_tiles + 15) / 16, 1, sm_count * ctas_per_sm), ENT, sizeof(EmitSmem), st>>>(
        b, n_tiles, sc.bitmap, sc.tile_base, ev_cap, ev_start, ev_mean, ev_stdv, sc.status, sc.tile_read0);
    sum_fixups_kernel<<<grid_cap(b.n_reads, 256, sm_count * 4), 256, 0, st>>>(b.n_reads, fixups, sc.counters);
    return 2;
}

}  // namespace sgpu
