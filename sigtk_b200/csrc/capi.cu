// capi.cu -- the C-ABI of the sigtk B200 hot path (include/sigtk_b200.h).
//
// Context = one device, its workspace in HBM, and n_slots pinned host slots.
// Host path per slot:  H2D (slot stream) -> kernels (context compute stream) -> D2H (slot stream),
// chained with events so that copies of one slot overlap the kernels of another while the
// kernels themselves are serialised (they share one scratch area).
#include "../../include/sigtk_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "kernels.cuh"

using namespace sgpu;

namespace {

struct DevOut {  // device-side outputs of one batch
    uint64_t* ev_off = nullptr;
    uint32_t* ev_start = nullptr;
    float* ev_mean = nullptr;
    float* ev_stdv = nullptr;
    float* pa = nullptr;
    float* stat = nullptr;
    double* ent = nullptr;  // [max_reads][3], allocated on first use
    uint32_t* jnn_cnt = nullptr;  // [max_reads], allocated on first use
    int32_t* jnn_seg = nullptr;   // [2 * jnn_seg_capacity]
    int32_t* prefix_pos = nullptr;   // [max_reads][4], allocated on first use
    float* prefix_stat = nullptr;    // [max_reads][6]
};

struct Slot {
    // pinned host input
    sgpu_batch_t batch{};
    uint64_t used = 0;
    // device input
    int16_t* d_samples = nullptr;
    uint64_t* d_read_off = nullptr;
    uint32_t* d_read_len = nullptr;
    float* d_offset = nullptr;
    float* d_unit = nullptr;
    DevOut dout;
    uint32_t* d_seq = nullptr;
    uint32_t* d_fix = nullptr;
    // pinned host output
    uint64_t* h_ev_off = nullptr;
    uint32_t* h_ev_start = nullptr;
    float* h_ev_mean = nullptr;
    float* h_ev_stdv = nullptr;
    float* h_pa = nullptr;
    float* h_stat = nullptr;
    double* h_ent = nullptr;
    uint32_t* h_jnn_cnt = nullptr;
    int32_t* h_jnn_seg = nullptr;
    int32_t* h_prefix_pos = nullptr;
    float* h_prefix_stat = nullptr;
    uint32_t* h_seq = nullptr;
    uint32_t* h_fix = nullptr;
    unsigned long long* h_counters = nullptr;  // [4]
    int* h_status = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t in_ready = nullptr, kernels_done = nullptr, small_back = nullptr;
    bool submitted = false;
    uint32_t want = 0;
    // svb-zd input (allocated on first use): the batch's compressed streams, decoded in HBM by svbzd.cu
    int mode = 0;                    // 0: empty, 1: decoded int16 records, 2: svb-zd streams
    uint8_t* h_comp = nullptr;       // pinned, comp_cap bytes
    uint8_t* d_comp = nullptr;       // comp_cap + 16 bytes
    uint64_t* h_comp_off = nullptr; uint64_t* d_comp_off = nullptr;
    uint32_t* h_comp_len = nullptr; uint32_t* d_comp_len = nullptr;
    uint64_t comp_used = 0, svb_blocks = 0;
};

}  // namespace

struct sgpu_ctx {
    int device = 0;
    int sm_count = 148;
    uint64_t max_samples = 0;
    uint32_t max_reads = 0;
    uint32_t n_slots = 0;
    uint32_t flags = 0;
    uint64_t ev_cap = 0;
    Scratch sc{};
    DevOut dev_out;              // outputs of the device-resident path
    uint32_t* dev_seq = nullptr; // seq_order / fixups of the device-resident path
    uint32_t* dev_fix = nullptr;
    Slot* slots = nullptr;
    SvbScratch svb{};            // workspace of the svb-zd decoder (allocated on first use)
    float* jnn_mom = nullptr;    // [max_reads][2] mean / stdv of the clamped signal (allocated on first use)
    int pore_rna004 = 0;         // SGPU_PARAM_PORE: jnnv2 parameters of `prefix` (misc.c:74-101 picks them from the header)
    uint32_t* ent_ovf = nullptr; // overflow histograms of ent_kernel (allocated on first use, kept all-zero)
    uint64_t comp_cap = 0;       // bytes of compressed input a slot can hold
    cudaStream_t compute = nullptr;
    uint64_t last_launches = 0;
    // optional per-stage CUDA-event timers (SGPU_F_STAGE_TIMERS)
    static constexpr int MAX_STAGES = 16;
    cudaEvent_t stage_ev[MAX_STAGES + 1] = {};
    const char* stage_name[MAX_STAGES] = {};
    uint32_t stage_launches[MAX_STAGES] = {};
    int n_stages = 0;
    cudaStream_t stage_stream = nullptr;
    char err[512] = {0};
};

namespace {

bool cuda_fail(sgpu_ctx* c, cudaError_t e, const char* what, int line) {
    if (e == cudaSuccess) return false;
    if (c) snprintf(c->err, sizeof c->err, "%s: %s (%s) at capi.cu:%d", what, cudaGetErrorName(e),
                    cudaGetErrorString(e), line);
    return true;
}
#define CU(call)                                                         \
    do {                                                                 \
        if (cuda_fail(ctx, (call), #call, __LINE__)) return SGPU_E_CUDA; \
    } while (0)

template <typename T>
cudaError_t dev_alloc(T** p, uint64_t count) {
    return cudaMalloc(reinterpret_cast<void**>(p), (size_t)(count ? count : 1) * sizeof(T));
}
template <typename T>
cudaError_t pin_alloc(T** p, uint64_t count) {
    return cudaHostAlloc(reinterpret_cast<void**>(p), (size_t)(count ? count : 1) * sizeof(T), cudaHostAllocDefault);
}

uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

__global__ void fill_u32_kernel(uint32_t* p, uint32_t n, uint32_t v) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = v;
}

__global__ void set_u64_kernel(unsigned long long* p, unsigned long long v) { *p = v; }

int alloc_dev_out(sgpu_ctx* ctx, DevOut& o) {
    CU(dev_alloc(&o.ev_off, (uint64_t)ctx->max_reads + 1));
    CU(dev_alloc(&o.ev_start, ctx->ev_cap));
    CU(dev_alloc(&o.ev_mean, ctx->ev_cap));
    CU(dev_alloc(&o.ev_stdv, ctx->ev_cap));
    CU(dev_alloc(&o.stat, (uint64_t)ctx->max_reads * 6));
    o.pa = nullptr;  // allocated on first use (4 B/sample)
    return SGPU_OK;
}

void free_dev_out(DevOut& o) {
    cudaFree(o.ev_off); cudaFree(o.ev_start); cudaFree(o.ev_mean); cudaFree(o.ev_stdv);
    cudaFree(o.pa); cudaFree(o.stat); cudaFree(o.ent); cudaFree(o.jnn_cnt); cudaFree(o.jnn_seg);
    cudaFree(o.prefix_pos); cudaFree(o.prefix_stat);
    o = DevOut{};
}

struct StageMarks {  // records a CUDA event between kernel groups when the context has stage timers
    sgpu_ctx* ctx;
    cudaStream_t st;
    bool on;
    uint64_t launches = 0;
    StageMarks(sgpu_ctx* c, cudaStream_t s) : ctx(c), st(s), on((c->flags & SGPU_F_STAGE_TIMERS) != 0) {
        c->n_stages = 0;
        c->stage_stream = s;
        if (on) cudaEventRecord(c->stage_ev[0], s);
    }
    void done(const char* name, int n) {
        launches += (uint64_t)n;
        if (!on || ctx->n_stages >= sgpu_ctx::MAX_STAGES) return;
        const int k = ctx->n_stages++;
        ctx->stage_name[k] = name;
        ctx->stage_launches[k] = (uint32_t)n;
        cudaEventRecord(ctx->stage_ev[k + 1], st);
    }
};

// The kernel sequence of one batch. No host synchronisation, no host<->device copies.
int alloc_svb_scratch(sgpu_ctx* ctx) {
    SvbScratch& w = ctx->svb;
    if (w.cnt) return SGPU_OK;
    w.max_blocks = svbzd_max_blocks(ctx->max_samples, ctx->max_reads);
    CU(dev_alloc(&w.cnt, ctx->max_reads));
    CU(dev_alloc(&w.base, (uint64_t)ctx->max_reads + 1));
    CU(dev_alloc(&w.blk_bytes, w.max_blocks));
    CU(dev_alloc(&w.blk_gpos, w.max_blocks + 1));
    CU(dev_alloc(&w.blk_sum, w.max_blocks));
    CU(dev_alloc(&w.blk_vpos, w.max_blocks + 1));
    CU(dev_alloc(&w.lane_sum, w.max_blocks * 32));
    CU(dev_alloc(&w.blk_read, w.max_blocks));
    return SGPU_OK;
}

int run_pipeline(sgpu_ctx* ctx, const DevBatch& b, uint32_t want, DevOut& o, uint32_t* d_seq, uint32_t* d_fix,
                 cudaStream_t st, const SvbBatch* svb = nullptr, int16_t* svb_out = nullptr) {
    Scratch& sc = ctx->sc;
    if (b.n_reads > ctx->max_reads || b.span > ctx->max_samples || (b.span & (SGPU_ALIGN - 1))) return SGPU_E_INVAL;
    CU(cudaMemsetAsync(sc.status, 0, sizeof(int), st));
    CU(cudaMemsetAsync(sc.counters, 0, 4 * sizeof(unsigned long long), st));
    StageMarks marks(ctx, st);
    if (b.n_reads == 0 || b.span == 0) {
        // nothing but empty records: every per-read output gets its defined value for an empty record (no event,
        // zero statistics / entropies / segments, "record not longer than the window" for prefix) instead of whatever
        // the previous batch left in the buffers
        const size_t nr = b.n_reads;
        CU(cudaMemsetAsync(o.ev_off, 0, (nr + 1) * sizeof(uint64_t), st));
        if (nr && (want & SGPU_WANT_STAT)) CU(cudaMemsetAsync(o.stat, 0, nr * 6 * sizeof(float), st));
        if (nr && (want & SGPU_WANT_ENT)) {
            if (!o.ent) CU(dev_alloc(&o.ent, (uint64_t)ctx->max_reads * 3));
            CU(cudaMemsetAsync(o.ent, 0, nr * 3 * sizeof(double), st));
        }
        if (nr && (want & SGPU_WANT_JNN)) {
            if (!o.jnn_cnt) {
                CU(dev_alloc(&o.jnn_cnt, ctx->max_reads));
                CU(dev_alloc(&o.jnn_seg, 2 * jnn_seg_capacity(ctx->max_samples, ctx->max_reads)));
            }
            CU(cudaMemsetAsync(o.jnn_cnt, 0, nr * sizeof(uint32_t), st));
        }
        if (nr && (want & SGPU_WANT_PREFIX)) {
            if (!o.prefix_pos) {
                CU(dev_alloc(&o.prefix_pos, (uint64_t)ctx->max_reads * 4));
                CU(dev_alloc(&o.prefix_stat, (uint64_t)ctx->max_reads * 6));
            }
            CU(cudaMemsetAsync(o.prefix_pos, 0xff, nr * 4 * sizeof(int32_t), st));   // (-1, -1, -1, -1)
            CU(cudaMemsetAsync(o.prefix_stat, 0, nr * 6 * sizeof(float), st));
        }
        if (nr && (want & SGPU_WANT_EVENTS)) {
            CU(cudaMemsetAsync(d_seq, 0, nr * sizeof(uint32_t), st));
            CU(cudaMemsetAsync(d_fix, 0, nr * sizeof(uint32_t), st));
        }
        ctx->last_launches = 0;
        return SGPU_OK;
    }
    if (svb) marks.done("svbzd_decode", launch_svbzd_decode(*svb, ctx->svb, ctx->sc, svb_out, ctx->sm_count, st));
    const bool force_generic = (ctx->flags & SGPU_F_FORCE_GENERIC) != 0;
    const bool events = (want & SGPU_WANT_EVENTS) != 0;
    if (want & SGPU_WANT_PA) {
        if (!o.pa) CU(dev_alloc(&o.pa, align_up(ctx->max_samples, SGPU_ALIGN)));
        // with events on the fast path the pA store is fused into walk_chunks_kernel
        if (!events || force_generic) marks.done("pa", launch_pa(b, o.pa, ctx->sm_count, st));
    }
    if (want & SGPU_WANT_STAT) {
        marks.done("stat_moments", launch_stat_moments(b, o.stat, ctx->sc.tune_stat_cta_min, ctx->sm_count, st));
        marks.done("stat_median", launch_stat_median(b, o.stat, ctx->sc.tune_stat_cta_min, ctx->sm_count, st));
    }
    if (want & SGPU_WANT_ENT) {
        if (!o.ent) CU(dev_alloc(&o.ent, (uint64_t)ctx->max_reads * 3));
        if (!ctx->ent_ovf) {
            const uint64_t words = ent_overflow_words(ctx->sm_count);
            CU(dev_alloc(&ctx->ent_ovf, words));
            CU(cudaMemsetAsync(ctx->ent_ovf, 0, (size_t)words * sizeof(uint32_t), st));
        }
        marks.done("ent", launch_ent(b, ctx->ent_ovf, o.ent, ctx->sm_count, st));
    }
    if (want & SGPU_WANT_JNN) {
        if (!o.jnn_cnt) {
            CU(dev_alloc(&o.jnn_cnt, ctx->max_reads));
            CU(dev_alloc(&o.jnn_seg, 2 * jnn_seg_capacity(ctx->max_samples, ctx->max_reads)));
        }
        if (!ctx->jnn_mom) CU(dev_alloc(&ctx->jnn_mom, (uint64_t)ctx->max_reads * 2));
        marks.done("jnn_moments", launch_jnn_moments(b, ctx->jnn_mom, ctx->sc.tune_stat_cta_min, ctx->sm_count, st));
        marks.done("jnn_walk", launch_jnn(b, ctx->jnn_mom, o.jnn_cnt, o.jnn_seg, ctx->sm_count, st));
    }
    if (want & SGPU_WANT_PREFIX) {
        if (!o.prefix_pos) {
            CU(dev_alloc(&o.prefix_pos, (uint64_t)ctx->max_reads * 4));
            CU(dev_alloc(&o.prefix_stat, (uint64_t)ctx->max_reads * 6));
        }
        marks.done("prefix", launch_prefix(b, ctx->pore_rna004, o.prefix_pos, o.prefix_stat, ctx->sm_count, st));
    }
    if (events) {
        const uint32_t n_tiles = fast_tiles_for(b.span);
        if (n_tiles > sc.max_tiles) return SGPU_E_INVAL;
        int n = launch_init_reads(b, sc, d_seq, d_fix, ctx->sm_count, st);
        if (force_generic) {
            CU(cudaMemsetAsync(sc.bitmap, 0, (size_t)n_tiles * (FAST_TILE / 32) * sizeof(uint32_t), st));
            marks.done("init", n);
        } else {
            marks.done("init", n);
            marks.done("walk_chunks", launch_walk(b, sc, (want & SGPU_WANT_PA) ? o.pa : nullptr, d_seq, ctx->sm_count, st));
            marks.done("long_jobs", launch_long_jobs(b, sc, d_seq, ctx->sm_count, st));
            marks.done("verify_chunks", launch_verify_chunks(b, sc, d_seq, d_fix, ctx->sm_count, st));
        }
        n = launch_build_seq_list(b, sc, d_seq, force_generic ? 1 : 0, ctx->sm_count, st);
        WorkList wl{sc.seq_list, sc.seq_sbase, sc.seq_count};
        n += launch_generic_detect(b, wl, sc, ctx->sm_count, st);
        marks.done("sequential_order_detect", n);
        marks.done("rank_events", launch_rank_events(b, sc, o.ev_off, ctx->sm_count, st));
        marks.done("emit_events", launch_fast_emit(b, sc, ctx->ev_cap, o.ev_start, o.ev_mean, o.ev_stdv, d_fix,
                                                  ctx->sm_count, st));
        marks.done("sequential_order_emit", launch_generic_emit(b, wl, sc, o.ev_off, ctx->ev_cap, o.ev_start, o.ev_mean,
                                                                o.ev_stdv, ctx->sm_count, st));
    }
    CU(cudaGetLastError());
    ctx->last_launches = marks.launches;
    return SGPU_OK;
}

int map_dev_status(int s) {
    if (s == SGPU_DEV_E_EVCAP) return SGPU_E_EVCAP;
    if (s == SGPU_DEV_E_SCRATCH) return SGPU_E_SCRATCH;
    if (s == SGPU_DEV_E_STREAM) return SGPU_E_STREAM;
    return SGPU_OK;
}

}  // namespace

extern "C" {

int sgpu_abi_version(void) { return SGPU_ABI_VERSION; }

int sgpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char* sgpu_strerror(int code) {
    switch (code) {
        case SGPU_OK: return "success";
        case SGPU_E_INVAL: return "invalid argument";
        case SGPU_E_CUDA: return "CUDA error (no usable B200 / kernel failure); there is no CPU fallback";
        case SGPU_E_NOMEM: return "out of memory";
        case SGPU_E_FULL: return "slot full";
        case SGPU_E_TOOBIG: return "read larger than the context's max_samples";
        case SGPU_E_STATE: return "call out of order";
        case SGPU_E_EVCAP: return "event capacity exceeded";
        case SGPU_E_SCRATCH: return "sequential-order scratch too small";
        case SGPU_E_STREAM: return "malformed svb-zd stream (length does not match its keys)";
        default: return "unknown error";
    }
}

const char* sgpu_last_error(const sgpu_ctx_t* ctx) { return ctx ? ctx->err : "no context"; }

int sgpu_create(sgpu_ctx_t** out, int device, uint64_t max_samples, uint32_t max_reads, uint32_t n_slots,
                uint32_t flags) {
    if (!out || max_samples == 0 || max_reads == 0 || (flags & ~15u)) return SGPU_E_INVAL;
    *out = nullptr;
    if (flags & SGPU_F_NO_HOST_SLOTS) n_slots = 0; else if (n_slots == 0) n_slots = 2;
    sgpu_ctx* ctx = new (std::nothrow) sgpu_ctx();
    if (!ctx) return SGPU_E_NOMEM;
    int rc = SGPU_OK;
    auto fail = [&](int code) { sgpu_destroy(ctx); return code; };
    {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || device < 0 || device >= ndev) {
            fprintf(stderr, "[sigtk_b200] no usable CUDA device %d (%s); this library has no CPU fallback\n", device,
                    e == cudaSuccess ? "index out of range" : cudaGetErrorString(e));
            delete ctx;
            return SGPU_E_CUDA;
        }
    }
#define CUC(call)                                                               \
    do {                                                                        \
        if (cuda_fail(ctx, (call), #call, __LINE__)) {                          \
            fprintf(stderr, "[sigtk_b200] %s\n", ctx->err);                     \
            return fail(SGPU_E_CUDA);                                           \
        }                                                                       \
    } while (0)
    CUC(cudaSetDevice(device));
    ctx->device = device;
    cudaDeviceProp prop;
    CUC(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    max_samples = align_up(max_samples, 64);
    ctx->max_samples = max_samples;
    ctx->max_reads = max_reads;
    ctx->n_slots = n_slots;
    ctx->flags = flags;
    // consecutive peaks are >= 3 samples apart (DNA; 5 for RNA) => <= n/3 + 2 events per read
    ctx->ev_cap = max_samples / 3 + 2ull * max_reads + 16;
    Scratch& sc = ctx->sc;
    // scratch of the sequential-order kernels (24 B/sample): everything for small contexts or when forced,
    // one eighth of the batch otherwise (reads that fail the fast path's checks are rare)
    // (SGPU_F_FULL_SEQ_SCRATCH: a context opened for one huge read must be able to redo that read in order)
    sc.gen_cap = (max_samples <= (64ull << 20) || (flags & (SGPU_F_FORCE_GENERIC | SGPU_F_FULL_SEQ_SCRATCH)))
                     ? max_samples : max_samples / 8;
    CUC(dev_alloc(&sc.Sinc, sc.gen_cap));
    CUC(dev_alloc(&sc.Qinc, sc.gen_cap));
    CUC(dev_alloc(&sc.t1, sc.gen_cap));
    CUC(dev_alloc(&sc.t2, sc.gen_cap));
    sc.max_tiles = fast_tiles_for(max_samples) + 1;
    sc.bitmap_words = (uint64_t)sc.max_tiles * (FAST_TILE / 32);
    CUC(dev_alloc(&sc.bitmap, sc.bitmap_words));
    sc.wk_slots = walk_state_slots(max_samples, max_reads);
    CUC(dev_alloc(&sc.wk_begin, sc.wk_slots * 8));
    CUC(dev_alloc(&sc.wk_end, sc.wk_slots * 8));
    CUC(dev_alloc(&sc.wk_cnt, max_reads));
    CUC(dev_alloc(&sc.wk_ibase, (uint64_t)max_reads + 1));
    sc.job_cap = walk_job_capacity(max_samples);
    CUC(dev_alloc(&sc.jobs, (uint64_t)sc.job_cap * 4));
    CUC(dev_alloc(&sc.job_count, 1));
    sc.tune_chunk_len = 0; sc.tune_warmup = 0; sc.tune_thr_long = 9.0f;  // events.c:46,53: threshold of the long detector
    sc.tune_stat_cta_min = STAT_CTA_MIN_DEFAULT;
    CUC(dev_alloc(&sc.tile_cnt, sc.max_tiles));
    CUC(dev_alloc(&sc.tile_read0, sc.max_tiles));
    CUC(dev_alloc(&sc.tile_base, (uint64_t)sc.max_tiles + 1));
    CUC(dev_alloc(&sc.wit_min, max_reads));
    CUC(dev_alloc(&sc.wit_max, max_reads));
    CUC(dev_alloc(&ctx->dev_seq, max_reads));
    CUC(dev_alloc(&ctx->dev_fix, max_reads));
    CUC(dev_alloc(&sc.seq_list, max_reads));
    CUC(dev_alloc(&sc.seq_sbase, (uint64_t)max_reads + 1));
    CUC(dev_alloc(&sc.seq_count, 1));
    CUC(dev_alloc(&sc.cursor, 1));
    {
        uint64_t most = sc.max_tiles > max_reads ? sc.max_tiles : max_reads;
        const uint64_t svb_blocks = svbzd_max_blocks(max_samples, max_reads);
        if (svb_blocks > most) most = svb_blocks;
        CUC(dev_alloc(&sc.scan_status, scan_tiles_for((uint32_t)most) + 1));
    }
    ctx->comp_cap = 2 * max_samples;
    if (flags & SGPU_F_STAGE_TIMERS)
        for (int k = 0; k <= sgpu_ctx::MAX_STAGES; k++) CUC(cudaEventCreate(&ctx->stage_ev[k]));
    CUC(dev_alloc(&sc.scan_ticket, 1));
    CUC(dev_alloc(&sc.status, 1));
    CUC(dev_alloc(&sc.counters, 4));
    CUC(cudaStreamCreateWithFlags(&ctx->compute, cudaStreamNonBlocking));
    rc = alloc_dev_out(ctx, ctx->dev_out);
    if (rc) return fail(rc);
    if (n_slots) {
        ctx->slots = new (std::nothrow) Slot[n_slots];
        if (!ctx->slots) return fail(SGPU_E_NOMEM);
        for (uint32_t s = 0; s < n_slots; s++) {
            Slot& sl = ctx->slots[s];
            CUC(pin_alloc(&sl.batch.samples, max_samples));
            CUC(pin_alloc(&sl.batch.read_off, (uint64_t)max_reads + 1));
            CUC(pin_alloc(&sl.batch.read_len, max_reads));
            CUC(pin_alloc(&sl.batch.offset_f, max_reads));
            CUC(pin_alloc(&sl.batch.raw_unit_f, max_reads));
            CUC(dev_alloc(&sl.d_samples, max_samples));
            CUC(dev_alloc(&sl.d_read_off, (uint64_t)max_reads + 1));
            CUC(dev_alloc(&sl.d_read_len, max_reads));
            CUC(dev_alloc(&sl.d_offset, max_reads));
            CUC(dev_alloc(&sl.d_unit, max_reads));
            CUC(dev_alloc(&sl.d_seq, max_reads));
            CUC(dev_alloc(&sl.d_fix, max_reads));
            // every slot gets its own outputs (its D2H overlaps the next slot's kernels)
            rc = alloc_dev_out(ctx, sl.dout);
            if (rc) return fail(rc);
            CUC(pin_alloc(&sl.h_ev_off, (uint64_t)max_reads + 1));
            CUC(pin_alloc(&sl.h_ev_start, ctx->ev_cap));
            CUC(pin_alloc(&sl.h_ev_mean, ctx->ev_cap));
            CUC(pin_alloc(&sl.h_ev_stdv, ctx->ev_cap));
            CUC(pin_alloc(&sl.h_stat, (uint64_t)max_reads * 6));
            CUC(pin_alloc(&sl.h_seq, max_reads));
            CUC(pin_alloc(&sl.h_fix, max_reads));
            CUC(pin_alloc(&sl.h_counters, 4));
            CUC(pin_alloc(&sl.h_status, 1));
            CUC(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
            CUC(cudaEventCreateWithFlags(&sl.in_ready, cudaEventDisableTiming));
            CUC(cudaEventCreateWithFlags(&sl.kernels_done, cudaEventDisableTiming));
            CUC(cudaEventCreateWithFlags(&sl.small_back, cudaEventDisableTiming));
            sl.batch.n_reads = 0;
            sl.batch.read_off[0] = 0;
        }
    }
#undef CUC
    *out = ctx;
    return SGPU_OK;
}

void sgpu_destroy(sgpu_ctx_t* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    Scratch& sc = ctx->sc;
    cudaFree(sc.Sinc); cudaFree(sc.Qinc); cudaFree(sc.t1); cudaFree(sc.t2); cudaFree(sc.bitmap);
    cudaFree(sc.wk_begin); cudaFree(sc.wk_end); cudaFree(sc.wk_cnt); cudaFree(sc.wk_ibase);
    cudaFree(sc.jobs); cudaFree(sc.job_count);
    cudaFree(sc.tile_cnt); cudaFree(sc.tile_base); cudaFree(sc.tile_read0);
    cudaFree(sc.wit_min); cudaFree(sc.wit_max); cudaFree(ctx->dev_seq); cudaFree(ctx->dev_fix);
    cudaFree(sc.seq_list); cudaFree(sc.seq_sbase); cudaFree(sc.seq_count); cudaFree(sc.cursor);
    cudaFree(sc.scan_status); cudaFree(sc.scan_ticket); cudaFree(sc.status); cudaFree(sc.counters);
    for (int k = 0; k <= sgpu_ctx::MAX_STAGES; k++) if (ctx->stage_ev[k]) cudaEventDestroy(ctx->stage_ev[k]);
    free_dev_out(ctx->dev_out);
    cudaFree(ctx->ent_ovf); cudaFree(ctx->jnn_mom);
    {
        SvbScratch& w = ctx->svb;
        cudaFree(w.cnt); cudaFree(w.base); cudaFree(w.blk_bytes); cudaFree(w.blk_gpos); cudaFree(w.blk_sum);
        cudaFree(w.blk_vpos); cudaFree(w.lane_sum); cudaFree(w.blk_read);
    }
    if (ctx->slots) {
        for (uint32_t s = 0; s < ctx->n_slots; s++) {
            Slot& sl = ctx->slots[s];
            cudaFreeHost(sl.batch.samples); cudaFreeHost(sl.batch.read_off); cudaFreeHost(sl.batch.read_len);
            cudaFreeHost(sl.batch.offset_f); cudaFreeHost(sl.batch.raw_unit_f);
            cudaFree(sl.d_samples); cudaFree(sl.d_read_off); cudaFree(sl.d_read_len); cudaFree(sl.d_offset);
            cudaFree(sl.d_unit); cudaFree(sl.d_seq); cudaFree(sl.d_fix);
            free_dev_out(sl.dout);
            cudaFreeHost(sl.h_ev_off); cudaFreeHost(sl.h_ev_start); cudaFreeHost(sl.h_ev_mean);
            cudaFreeHost(sl.h_ev_stdv); cudaFreeHost(sl.h_pa); cudaFreeHost(sl.h_stat); cudaFreeHost(sl.h_ent); cudaFreeHost(sl.h_jnn_cnt); cudaFreeHost(sl.h_jnn_seg); cudaFreeHost(sl.h_prefix_pos); cudaFreeHost(sl.h_prefix_stat); cudaFreeHost(sl.h_seq);
            cudaFreeHost(sl.h_fix); cudaFreeHost(sl.h_counters); cudaFreeHost(sl.h_status);
            cudaFreeHost(sl.h_comp); cudaFreeHost(sl.h_comp_off); cudaFreeHost(sl.h_comp_len);
            cudaFree(sl.d_comp); cudaFree(sl.d_comp_off); cudaFree(sl.d_comp_len);
            if (sl.stream) cudaStreamDestroy(sl.stream);
            if (sl.in_ready) cudaEventDestroy(sl.in_ready);
            if (sl.kernels_done) cudaEventDestroy(sl.kernels_done);
            if (sl.small_back) cudaEventDestroy(sl.small_back);
        }
        delete[] ctx->slots;
    }
    if (ctx->compute) cudaStreamDestroy(ctx->compute);
    delete ctx;
}

int sgpu_slot_batch(sgpu_ctx_t* ctx, uint32_t slot, sgpu_batch_t** out) {
    if (!ctx || !out || slot >= ctx->n_slots) return SGPU_E_INVAL;
    *out = &ctx->slots[slot].batch;
    return SGPU_OK;
}

int sgpu_slot_reset(sgpu_ctx_t* ctx, uint32_t slot, uint32_t rna) {
    if (!ctx || slot >= ctx->n_slots) return SGPU_E_INVAL;
    Slot& sl = ctx->slots[slot];
    if (sl.submitted) return SGPU_E_STATE;
    sl.batch.n_reads = 0;
    sl.batch.rna = rna ? 1u : 0u;
    sl.batch.read_off[0] = 0;
    sl.used = 0;
    sl.mode = 0;
    sl.comp_used = 0;
    sl.svb_blocks = 0;
    return SGPU_OK;
}

int64_t sgpu_slot_add_read(sgpu_ctx_t* ctx, uint32_t slot, const int16_t* raw, uint64_t n, double digitisation,
                           double offset, double range) {
    if (!ctx || slot >= ctx->n_slots || (!raw && n)) return SGPU_E_INVAL;
    Slot& sl = ctx->slots[slot];
    if (sl.submitted) return SGPU_E_STATE;
    if (sl.mode == 2) return SGPU_E_STATE;  // a batch holds decoded records or svb-zd streams, not both
    if (n >= (1ull << 31)) return SGPU_E_TOOBIG;  // the reference itself narrows to int32_t (misc.c:20)
    const uint64_t need = align_up(n, SGPU_ALIGN);
    if (need > ctx->max_samples) return SGPU_E_TOOBIG;
    sgpu_batch_t& b = sl.batch;
    if (b.n_reads >= ctx->max_reads || sl.used + need > ctx->max_samples) return SGPU_E_FULL;
    sl.mode = 1;
    const uint32_t r = b.n_reads;
    memcpy(b.samples + sl.used, raw, (size_t)n * sizeof(int16_t));
    b.read_off[r] = sl.used;
    b.read_len[r] = (uint32_t)n;
    // misc.c:17-19,26: narrow the three doubles to float first, then ONE float division
    const float range_f = (float)range, dig_f = (float)digitisation, off_f = (float)offset;
    volatile float unit = range_f / dig_f;
    b.offset_f[r] = off_f;
    b.raw_unit_f[r] = unit;
    sl.used += need;
    b.n_reads = r + 1;
    b.read_off[r + 1] = sl.used;
    return (int64_t)r;
}

int64_t sgpu_slot_add_read_svbzd(sgpu_ctx_t* ctx, uint32_t slot, const uint8_t* stream, uint64_t n_bytes,
                                 double digitisation, double offset, double range) {
    if (!ctx || slot >= ctx->n_slots || !stream) return SGPU_E_INVAL;
    Slot& sl = ctx->slots[slot];
    if (sl.submitted || sl.mode == 1) return SGPU_E_STATE;
    if (n_bytes != 0 && n_bytes < 4) return SGPU_E_STREAM;
    uint32_t count = 0;
    if (n_bytes) memcpy(&count, stream, 4);  // slow5_press.c:1093: the original length leads the stream
    const uint64_t n = count;                // (a record without signal stores no stream at all: n_bytes == 0)
    if (n_bytes && (4u + (n + 3u) / 4u + n > n_bytes || n_bytes > 4u + (n + 3u) / 4u + 4u * n)) return SGPU_E_STREAM;
    if (n >= (1ull << 31) || n_bytes > 0xffffffffull) return SGPU_E_TOOBIG;
    const uint64_t need = align_up(n, SGPU_ALIGN), cneed = align_up(n_bytes, 16);
    if (need > ctx->max_samples || cneed > ctx->comp_cap) return SGPU_E_TOOBIG;
    sgpu_batch_t& b = sl.batch;
    if (b.n_reads >= ctx->max_reads || sl.used + need > ctx->max_samples || sl.comp_used + cneed > ctx->comp_cap)
        return SGPU_E_FULL;
    if (!sl.h_comp) {
        CU(cudaSetDevice(ctx->device));
        const int rc = alloc_svb_scratch(ctx);
        if (rc) return rc;
        CU(pin_alloc(&sl.h_comp, ctx->comp_cap));
        CU(dev_alloc(&sl.d_comp, ctx->comp_cap + 16));
        CU(pin_alloc(&sl.h_comp_off, ctx->max_reads));
        CU(dev_alloc(&sl.d_comp_off, ctx->max_reads));
        CU(pin_alloc(&sl.h_comp_len, ctx->max_reads));
        CU(dev_alloc(&sl.d_comp_len, ctx->max_reads));
    }
    sl.mode = 2;
    const uint32_t r = b.n_reads;
    if (n_bytes) memcpy(sl.h_comp + sl.comp_used, stream, (size_t)n_bytes);
    sl.h_comp_off[r] = sl.comp_used;
    sl.h_comp_len[r] = (uint32_t)n_bytes;
    sl.comp_used += cneed;
    sl.svb_blocks += svbzd_blocks_of(n);
    b.read_off[r] = sl.used;
    b.read_len[r] = (uint32_t)n;
    const float range_f = (float)range, dig_f = (float)digitisation, off_f = (float)offset;  // misc.c:17-19,26
    volatile float unit = range_f / dig_f;
    b.offset_f[r] = off_f;
    b.raw_unit_f[r] = unit;
    sl.used += need;
    b.n_reads = r + 1;
    b.read_off[r + 1] = sl.used;
    return (int64_t)r;
}

int sgpu_submit(sgpu_ctx_t* ctx, uint32_t slot, uint32_t want) {
    if (!ctx || slot >= ctx->n_slots || (want & ~63u) || want == 0) return SGPU_E_INVAL;
    Slot& sl = ctx->slots[slot];
    if (sl.submitted) return SGPU_E_STATE;
    CU(cudaSetDevice(ctx->device));
    const sgpu_batch_t& hb = sl.batch;
    const uint32_t nr = hb.n_reads;
    cudaStream_t st = sl.stream;
    SvbBatch svb{};
    const bool compressed = sl.mode == 2;
    if (compressed) {
        CU(cudaMemcpyAsync(sl.d_comp, sl.h_comp, (size_t)sl.comp_used, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(sl.d_comp_off, sl.h_comp_off, (size_t)nr * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(sl.d_comp_len, sl.h_comp_len, (size_t)nr * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        svb = SvbBatch{sl.d_comp, sl.comp_used, sl.d_comp_off, sl.d_comp_len, sl.d_read_off, sl.d_read_len, nr,
                       sl.svb_blocks};
    } else {
        CU(cudaMemcpyAsync(sl.d_samples, hb.samples, (size_t)sl.used * sizeof(int16_t), cudaMemcpyHostToDevice, st));
    }
    CU(cudaMemcpyAsync(sl.d_read_off, hb.read_off, ((size_t)nr + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(sl.d_read_len, hb.read_len, (size_t)nr * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(sl.d_offset, hb.offset_f, (size_t)nr * sizeof(float), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(sl.d_unit, hb.raw_unit_f, (size_t)nr * sizeof(float), cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(sl.in_ready, st));
    CU(cudaStreamWaitEvent(ctx->compute, sl.in_ready, 0));
    DevBatch b{sl.d_samples, sl.d_read_off, sl.d_read_len, sl.d_offset, sl.d_unit, nr, (int)hb.rna, sl.used};
    if ((want & SGPU_WANT_PA) && !sl.h_pa) CU(pin_alloc(&sl.h_pa, ctx->max_samples));
    if ((want & SGPU_WANT_ENT) && !sl.h_ent) CU(pin_alloc(&sl.h_ent, (uint64_t)ctx->max_reads * 3));
    if ((want & SGPU_WANT_JNN) && !sl.h_jnn_cnt) {
        CU(pin_alloc(&sl.h_jnn_cnt, ctx->max_reads));
        CU(pin_alloc(&sl.h_jnn_seg, 2 * jnn_seg_capacity(ctx->max_samples, ctx->max_reads)));
    }
    if ((want & SGPU_WANT_PREFIX) && !sl.h_prefix_pos) {
        CU(pin_alloc(&sl.h_prefix_pos, (uint64_t)ctx->max_reads * 4));
        CU(pin_alloc(&sl.h_prefix_stat, (uint64_t)ctx->max_reads * 6));
    }
    const int rc = run_pipeline(ctx, b, want, sl.dout, sl.d_seq, sl.d_fix, ctx->compute, compressed ? &svb : nullptr,
                                sl.d_samples);
    if (rc) return rc;
    // small results come back right away; the big arrays are sized by n_events in sgpu_wait
    CU(cudaMemcpyAsync(sl.h_counters, ctx->sc.counters, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                       ctx->compute));
    CU(cudaMemcpyAsync(sl.h_status, ctx->sc.status, sizeof(int), cudaMemcpyDeviceToHost, ctx->compute));
    CU(cudaEventRecord(sl.kernels_done, ctx->compute));
    CU(cudaStreamWaitEvent(st, sl.kernels_done, 0));
    if (want & SGPU_WANT_EVENTS) {
        CU(cudaMemcpyAsync(sl.h_ev_off, sl.dout.ev_off, ((size_t)nr + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(sl.h_seq, sl.d_seq, (size_t)nr * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(sl.h_fix, sl.d_fix, (size_t)nr * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    }
    if (want & SGPU_WANT_STAT)
        CU(cudaMemcpyAsync(sl.h_stat, sl.dout.stat, (size_t)nr * 6 * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (want & SGPU_WANT_ENT)
        CU(cudaMemcpyAsync(sl.h_ent, sl.dout.ent, (size_t)nr * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (want & SGPU_WANT_JNN) {
        CU(cudaMemcpyAsync(sl.h_jnn_cnt, sl.dout.jnn_cnt, (size_t)nr * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(sl.h_jnn_seg, sl.dout.jnn_seg, (size_t)2 * ((sl.used >> 5) + nr + 1) * sizeof(int32_t),
                           cudaMemcpyDeviceToHost, st));
    }
    if (want & SGPU_WANT_PREFIX) {
        CU(cudaMemcpyAsync(sl.h_prefix_pos, sl.dout.prefix_pos, (size_t)nr * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(sl.h_prefix_stat, sl.dout.prefix_stat, (size_t)nr * 6 * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    if (want & SGPU_WANT_PA)
        CU(cudaMemcpyAsync(sl.h_pa, sl.dout.pa, (size_t)sl.used * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(sl.small_back, st));
    sl.want = want;
    sl.submitted = true;
    return SGPU_OK;
}

int sgpu_wait(sgpu_ctx_t* ctx, uint32_t slot, sgpu_result_t* out) {
    if (!ctx || slot >= ctx->n_slots || !out) return SGPU_E_INVAL;
    Slot& sl = ctx->slots[slot];
    if (!sl.submitted) return SGPU_E_STATE;
    CU(cudaSetDevice(ctx->device));
    sl.submitted = false;
    CU(cudaEventSynchronize(sl.small_back));
    memset(out, 0, sizeof *out);
    const int dev_rc = map_dev_status(*sl.h_status);
    if (dev_rc) return dev_rc;
    const uint32_t nr = sl.batch.n_reads;
    if (sl.want & SGPU_WANT_EVENTS) {
        const uint64_t ne = nr ? sl.h_ev_off[nr] : 0;
        if (ne > ctx->ev_cap) return SGPU_E_EVCAP;
        cudaStream_t st = sl.stream;
        CU(cudaMemcpyAsync(sl.h_ev_start, sl.dout.ev_start, (size_t)ne * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(sl.h_ev_mean, sl.dout.ev_mean, (size_t)ne * sizeof(float), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(sl.h_ev_stdv, sl.dout.ev_stdv, (size_t)ne * sizeof(float), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        out->ev_off = sl.h_ev_off;
        out->ev_start = sl.h_ev_start;
        out->ev_mean = sl.h_ev_mean;
        out->ev_stdv = sl.h_ev_stdv;
        out->seq_order = sl.h_seq;
        out->fixups = sl.h_fix;
        out->n_events = ne;
    }
    if (sl.want & SGPU_WANT_STAT) out->stat = sl.h_stat;
    if (sl.want & SGPU_WANT_PA) out->pa = sl.h_pa;
    if (sl.want & SGPU_WANT_ENT) out->ent = sl.h_ent;
    if (sl.want & SGPU_WANT_JNN) { out->jnn_cnt = sl.h_jnn_cnt; out->jnn_seg = sl.h_jnn_seg; }
    if (sl.want & SGPU_WANT_PREFIX) { out->prefix_pos = sl.h_prefix_pos; out->prefix_stat = sl.h_prefix_stat; }
    return SGPU_OK;
}

int sgpu_run_device(sgpu_ctx_t* ctx, const sgpu_dev_batch_t* batch, uint32_t want, void* stream,
                    sgpu_result_t* out) {
    if (!ctx || !batch || !out || (want & ~63u) || want == 0) return SGPU_E_INVAL;
    CU(cudaSetDevice(ctx->device));
    DevBatch b{batch->samples, batch->read_off, batch->read_len, batch->offset_f, batch->raw_unit_f,
               batch->n_reads, (int)batch->rna, batch->span};
    const int rc = run_pipeline(ctx, b, want, ctx->dev_out, ctx->dev_seq, ctx->dev_fix,
                                reinterpret_cast<cudaStream_t>(stream));
    if (rc) return rc;
    memset(out, 0, sizeof *out);
    out->ev_off = ctx->dev_out.ev_off;
    out->ev_start = ctx->dev_out.ev_start;
    out->ev_mean = ctx->dev_out.ev_mean;
    out->ev_stdv = ctx->dev_out.ev_stdv;
    out->pa = (want & SGPU_WANT_PA) ? ctx->dev_out.pa : nullptr;
    out->stat = ctx->dev_out.stat;
    out->ent = (want & SGPU_WANT_ENT) ? ctx->dev_out.ent : nullptr;
    out->jnn_cnt = (want & SGPU_WANT_JNN) ? ctx->dev_out.jnn_cnt : nullptr;
    out->jnn_seg = (want & SGPU_WANT_JNN) ? ctx->dev_out.jnn_seg : nullptr;
    out->prefix_pos = (want & SGPU_WANT_PREFIX) ? ctx->dev_out.prefix_pos : nullptr;
    out->prefix_stat = (want & SGPU_WANT_PREFIX) ? ctx->dev_out.prefix_stat : nullptr;
    out->seq_order = ctx->dev_seq;
    out->fixups = ctx->dev_fix;
    return SGPU_OK;
}

int sgpu_decode_svbzd_device(sgpu_ctx_t* ctx, const sgpu_svb_dev_batch_t* batch, int16_t* samples_out, void* stream) {
    if (!ctx || !batch || !samples_out) return SGPU_E_INVAL;
    if (batch->n_reads > ctx->max_reads || batch->n_blocks > svbzd_max_blocks(ctx->max_samples, ctx->max_reads))
        return SGPU_E_INVAL;
    CU(cudaSetDevice(ctx->device));
    const int arc = alloc_svb_scratch(ctx);
    if (arc) return arc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CU(cudaMemsetAsync(ctx->sc.status, 0, sizeof(int), st));
    SvbBatch s{batch->bytes, batch->n_bytes, batch->comp_off, batch->comp_len, batch->read_off, batch->read_len,
               batch->n_reads, batch->n_blocks};
    ctx->last_launches = (uint64_t)launch_svbzd_decode(s, ctx->svb, ctx->sc, samples_out, ctx->sm_count, st);
    CU(cudaGetLastError());
    return SGPU_OK;
}

int sgpu_counters(sgpu_ctx_t* ctx, sgpu_counters_t* out) {
    if (!ctx || !out) return SGPU_E_INVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaDeviceSynchronize());
    unsigned long long c[4];
    int status = 0;
    CU(cudaMemcpy(c, ctx->sc.counters, sizeof c, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&status, ctx->sc.status, sizeof status, cudaMemcpyDeviceToHost));
    out->n_events = c[0];
    out->n_seq_order_reads = c[1];
    out->n_fixups = c[2];
    out->n_kernel_launches = ctx->last_launches;
    out->status = map_dev_status(status);
    out->n_long_jobs = c[3];
    return SGPU_OK;
}

int sgpu_set_param(sgpu_ctx_t* ctx, int key, double value) {
    if (!ctx) return SGPU_E_INVAL;
    Scratch& sc = ctx->sc;
    switch (key) {
        case SGPU_PARAM_CHUNK_LEN:
            if (value < 0 || value > (double)(1u << 20)) return SGPU_E_INVAL;
            sc.tune_chunk_len = (uint32_t)value;
            return SGPU_OK;
        case SGPU_PARAM_WARMUP:
            if (value < 0 || value > 4096) return SGPU_E_INVAL;
            sc.tune_warmup = (uint32_t)value;
            return SGPU_OK;
        case SGPU_PARAM_PORE:
            if (value != 0 && value != 1) return SGPU_E_INVAL;
            ctx->pore_rna004 = (int)value;
            return SGPU_OK;
        case SGPU_PARAM_STAT_CTA_MIN:
            if (value < 0 || value > 4294967295.0) return SGPU_E_INVAL;
            sc.tune_stat_cta_min = (uint32_t)value;
            return SGPU_OK;
        case SGPU_PARAM_THR_LONG:
            if (!(value > 0 && value < 1e6)) return SGPU_E_INVAL;
            sc.tune_thr_long = (float)value;
            return SGPU_OK;
        default: return SGPU_E_INVAL;
    }
}

int sgpu_stage_times(sgpu_ctx_t* ctx, sgpu_stage_time_t* out, uint32_t cap) {
    if (!ctx || (!out && cap)) return SGPU_E_INVAL;
    if (!(ctx->flags & SGPU_F_STAGE_TIMERS)) return SGPU_E_STATE;
    CU(cudaSetDevice(ctx->device));
    if (ctx->n_stages == 0) return 0;
    CU(cudaEventSynchronize(ctx->stage_ev[ctx->n_stages]));
    int n = 0;
    for (int k = 0; k < ctx->n_stages && (uint32_t)k < cap; k++, n++) {
        float ms = 0.0f;
        CU(cudaEventElapsedTime(&ms, ctx->stage_ev[k], ctx->stage_ev[k + 1]));
        out[k].name = ctx->stage_name[k];
        out[k].ms = ms;
        out[k].launches = ctx->stage_launches[k];
    }
    return n;
}

int sgpu_memcpy_d2h(sgpu_ctx_t* ctx, void* dst, const void* src, uint64_t bytes) {
    if (!ctx || (!dst && bytes) || (!src && bytes)) return SGPU_E_INVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaDeviceSynchronize());
    if (bytes) CU(cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost));
    return SGPU_OK;
}

}  // extern "C"
