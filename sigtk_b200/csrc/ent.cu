// ent.cu -- `sigtk ent`: per-read Shannon entropy of the raw samples, of their zig-zag deltas and of the
// deltas' byte planes (reference src/ent.c:25-51 `entropy`, 56-65 `zigzag_delta_encode`, 108-151 the per-record
// loop of `entmain`). SURVEY 8f rank 4.
//
// One CTA per read, four histograms from ONE pass over the samples:
//   raw   : 65,536 bins keyed by (uint16)raw[i],                    i in [0, n)
//   delta : 65,536 bins keyed by (uint16)zigzag32(raw[i]-raw[i-1]), i in [0, n-1), raw[-1] = 0  (ent.c:118-131)
//   hi/lo : 256 bins each, the two bytes of the delta key (ent.c:139-147): not counted per sample but folded
//           from the delta histogram's nonzero bins while it is swept (they are its marginals)
// The two wide histograms live in SHARED memory as windows of ENT_WIN bins: raw keys within +-ENT_WIN/2 of the
// read's first sample, delta keys below ENT_WIN (nanopore signals span ~1,000 ADC units, their deltas a few
// hundred). A key outside its window goes to the CTA's 65,536-bin overflow histogram in HBM (L2 atomics) -- any
// int16 input is handled, only slower. The entropy sweep visits the bins between the smallest and largest key
// seen, clears them on the way (no per-read memset), and subtracts the terms p*log2(p) in ASCENDING KEY ORDER
// on one thread, because the reference's `ent -= p*log2(p)` loop (ent.c:39-45) is order dependent in the last
// bits. Counts are exact; log2 is CUDA's (<= 1 ulp, like glibc's), so the doubles agree with the reference to
// ~1e-15 and its "%f" text is identical.
#include "kernels.cuh"

namespace sgpu {

constexpr int ENT_WIN = 4096;      // bins per shared-memory window
constexpr int ENT_THREADS = 256;
constexpr int ENT_CTAS_PER_SM = 4; // 35 KB of shared memory per CTA

__device__ __forceinline__ uint32_t zigzag16(int32_t cur, int32_t prev) {
    const int32_t d = cur - prev;                             // ent.c:62
    return ((uint32_t)((d + d) ^ (d >> 31))) & 0xffffu;       // ent.c:56-58, narrowed to int16 at ent.c:128
}

struct EntShared {
    uint32_t raw[ENT_WIN];
    uint32_t dlt[ENT_WIN];
    uint32_t hi[256];
    uint32_t lo[256];
    double   term[ENT_THREADS];
    uint32_t mask[ENT_THREADS / 32];
    uint32_t kmin[2], kmax[2];  // smallest / largest key seen: [0] raw, [1] delta
    uint32_t overflow[2];       // a key left its window
};

// Entropy of one histogram over the keys [kmin, kmax]; `count(k)` returns the bin and clears it.
// All threads of the CTA call it; the result is valid on thread 0.
template <typename F>
__device__ double entropy_sweep(EntShared& s, uint32_t kmin, uint32_t kmax, double len, F count) {
    double ent = 0.0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = kmin; base <= kmax; base += ENT_THREADS) {
        const uint32_t k = base + threadIdx.x;
        const uint32_t c = (k <= kmax) ? count(k) : 0u;
        double t = 0.0;
        if (c) {
            const double p = __ddiv_rn((double)c, len);       // ent.c:41
            t = __dmul_rn(p, log2(p));                        // ent.c:42
        }
        // the nonzero terms of the chunk, packed in key order, so that the ordered subtraction below is one
        // dependent FP64 operation per term and nothing else
        const uint32_t m = __ballot_sync(0xffffffffu, c != 0);
        if (lane == 0) s.mask[warp] = m;
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < ENT_THREADS / 32; w++) {
            const uint32_t pc = __popc(s.mask[w]);
            if (w < warp) before += pc;
            total += pc;
        }
        if (c) s.term[before + __popc(m & ((1u << lane) - 1u))] = t;
        __syncthreads();
        if (threadIdx.x == 0)
            for (uint32_t q = 0; q < total; q++) ent = __dsub_rn(ent, s.term[q]);  // ent.c:42, ascending key order
        __syncthreads();
    }
    return ent;
}

__global__ void __launch_bounds__(ENT_THREADS, ENT_CTAS_PER_SM)
ent_kernel(DevBatch b, uint32_t* __restrict__ ovf_all, double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EntShared& s = *reinterpret_cast<EntShared*>(smem_raw);
    uint32_t* ovf_raw = ovf_all + (size_t)blockIdx.x * 2 * 65536;  // kept all-zero between reads
    uint32_t* ovf_dlt = ovf_raw + 65536;
    const int tid = threadIdx.x;
    for (int k = tid; k < ENT_WIN; k += ENT_THREADS) { s.raw[k] = 0; s.dlt[k] = 0; }
    if (tid < 256) { s.hi[tid] = 0; s.lo[tid] = 0; }
    __syncthreads();

    for (uint32_t r = blockIdx.x; r < b.n_reads; r += gridDim.x) {
        const int16_t* __restrict__ raw = b.samples + b.read_off[r];  // 16-byte aligned
        const uint32_t n = b.read_len[r];
        if (n == 0) {  // the reference does not survive an empty record (len-1 wraps, ent.c:131); defined here as zeros
            if (tid == 0) { out[(size_t)r * 3] = 0.0; out[(size_t)r * 3 + 1] = 0.0; out[(size_t)r * 3 + 2] = 0.0; }
            continue;
        }
        if (tid == 0) {
            s.kmin[0] = s.kmin[1] = 0xffffu; s.kmax[0] = s.kmax[1] = 0u;
            s.overflow[0] = s.overflow[1] = 0u;
        }
        __syncthreads();
        const uint32_t rbase = ((uint32_t)(uint16_t)raw[0] - ENT_WIN / 2) & 0xffffu;
        uint32_t rmin = 0xffffu, rmax = 0u, dmin = 0xffffu, dmax = 0u;
        bool rovf = false, dovf = false;
        const uint32_t nd = n - 1;  // deltas counted (ent.c:131)
        // 8 samples per thread and step: one 128-bit load + the sample before it; the next step's are in flight
        // while this one is counted
        auto load8 = [&](uint32_t i0, uint4& q, int32_t& prev) {
            if (i0 + 8 <= n) {
                q = __ldg(reinterpret_cast<const uint4*>(raw + i0));
            } else {
                int16_t t[8];
                for (int j = 0; j < 8; j++) t[j] = (i0 + j < n) ? raw[i0 + j] : (int16_t)0;
                q = *reinterpret_cast<uint4*>(t);
            }
            prev = i0 ? (int32_t)raw[i0 - 1] : 0;  // ent.c:124: prev starts at 0
        };
        auto count1 = [&](int32_t val, int32_t prev, bool with_delta) {
            const uint32_t key = (uint32_t)val & 0xffffu;
            rmin = min(rmin, key); rmax = max(rmax, key);
            const uint32_t d = (key - rbase) & 0xffffu;
            if (d < ENT_WIN) atomicAdd(&s.raw[d], 1u);
            else { atomicAdd(&ovf_raw[key], 1u); rovf = true; }
            if (with_delta) {
                const uint32_t z = zigzag16(val, prev);
                dmin = min(dmin, z); dmax = max(dmax, z);
                if (z < ENT_WIN) atomicAdd(&s.dlt[z], 1u);
                else { atomicAdd(&ovf_dlt[z], 1u); dovf = true; }
            }
        };
        uint4 cur = make_uint4(0, 0, 0, 0), nxt = make_uint4(0, 0, 0, 0);
        int32_t prev_c = 0, prev_n = 0;
        if ((uint32_t)tid * 8 < n) load8((uint32_t)tid * 8, cur, prev_c);
        for (uint32_t i0 = (uint32_t)tid * 8; i0 < n; i0 += ENT_THREADS * 8) {
            const uint32_t i1 = i0 + ENT_THREADS * 8;
            if (i1 < n) load8(i1, nxt, prev_n);
            const uint32_t wd[4] = {cur.x, cur.y, cur.z, cur.w};
            int32_t prev = prev_c;
            if (i0 + 8 <= nd) {   // every sample of the step has a delta (all but the read's last sample do)
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int32_t val = (int32_t)(int16_t)(wd[j >> 1] >> ((j & 1) * 16));
                    count1(val, prev, true);
                    prev = val;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const uint32_t i = i0 + j;
                    if (i >= n) break;
                    const int32_t val = (int32_t)(int16_t)(wd[j >> 1] >> ((j & 1) * 16));
                    count1(val, prev, i < nd);
                    prev = val;
                }
            }
            cur = nxt;
            prev_c = prev_n;
        }
        // CTA-wide key ranges
        for (int o = 16; o; o >>= 1) {
            rmin = min(rmin, __shfl_xor_sync(0xffffffffu, rmin, o)); rmax = max(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
            dmin = min(dmin, __shfl_xor_sync(0xffffffffu, dmin, o)); dmax = max(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
        }
        const bool rov_w = __any_sync(0xffffffffu, rovf), dov_w = __any_sync(0xffffffffu, dovf);
        if ((tid & 31) == 0) {
            atomicMin(&s.kmin[0], rmin); atomicMax(&s.kmax[0], rmax);
            atomicMin(&s.kmin[1], dmin); atomicMax(&s.kmax[1], dmax);
            if (rov_w) s.overflow[0] = 1u;
            if (dov_w) s.overflow[1] = 1u;
        }
        __threadfence();  // overflow bins (L2 atomics of other threads) are read below
        __syncthreads();
        const uint32_t k0 = s.kmin[0], k1 = s.kmax[0], z0 = s.kmin[1], z1 = s.kmax[1];
        const bool rov = s.overflow[0] != 0, dov = s.overflow[1] != 0;

        const double e_raw = entropy_sweep(s, k0, k1, (double)n, [&](uint32_t k) -> uint32_t {
            const uint32_t d = (k - rbase) & 0xffffu;
            uint32_t c;
            if (d < ENT_WIN) { c = s.raw[d]; s.raw[d] = 0; }
            else if (rov) { c = __ldcg(&ovf_raw[k]); if (c) ovf_raw[k] = 0; }
            else c = 0;
            return c;
        });
        double e_dlt = 0.0, e_hi = 0.0, e_lo = 0.0;
        if (nd) {
            const double len = (double)nd;
            e_dlt = entropy_sweep(s, z0, z1, len, [&](uint32_t k) -> uint32_t {
                uint32_t c;
                if (k < ENT_WIN) { c = s.dlt[k]; s.dlt[k] = 0; }
                else if (dov) { c = __ldcg(&ovf_dlt[k]); if (c) ovf_dlt[k] = 0; }
                else c = 0;
                if (c) {  // the byte planes' histograms are the marginals of this one (ent.c:144-145)
                    atomicAdd(&s.hi[k >> 8], c);
                    atomicAdd(&s.lo[k & 255u], c);
                }
                return c;
            });
            e_hi = entropy_sweep(s, 0u, 255u, len, [&](uint32_t k) -> uint32_t { const uint32_t c = s.hi[k]; s.hi[k] = 0; return c; });
            e_lo = entropy_sweep(s, 0u, 255u, len, [&](uint32_t k) -> uint32_t { const uint32_t c = s.lo[k]; s.lo[k] = 0; return c; });
        }
        if (tid == 0) {
            out[(size_t)r * 3] = e_raw;
            out[(size_t)r * 3 + 1] = e_dlt;
            out[(size_t)r * 3 + 2] = __dadd_rn(e_hi, e_lo);  // ent.c:147
        }
        __threadfence();  // cleared overflow bins are visible before the next read's atomics
        __syncthreads();
    }
}

uint32_t ent_grid(int sm_count) { return (uint32_t)(sm_count * ENT_CTAS_PER_SM); }
uint64_t ent_overflow_words(int sm_count) { return (uint64_t)ent_grid(sm_count) * 2 * 65536; }

// `overflow` : ent_overflow_words() zeroed words owned by the context. `out`: [n_reads][3] doubles.
int launch_ent(const DevBatch& b, uint32_t* overflow, double* out, int sm_count, cudaStream_t st) {
    if (b.n_reads == 0) return 0;
    cudaFuncSetAttribute(ent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EntShared));  // per device
    uint32_t g = ent_grid(sm_count);
    if (g > b.n_reads) g = b.n_reads;
    ent_kernel<<<g, ENT_THREADS, sizeof(EntShared), st>>>(b, overflow, out);
    return 1;
}

}  // namespace sgpu
