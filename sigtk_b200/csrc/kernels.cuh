// kernels.cuh -- internal interfaces between the kernel translation units and capi.cu
#pragma once
#include "common.cuh"

namespace sgpu {

// device-side status codes (mapped to SGPU_E_* by capi.cu)
enum { SGPU_DEV_OK = 0, SGPU_DEV_E_EVCAP = 1, SGPU_DEV_E_SCRATCH = 2, SGPU_DEV_E_STREAM = 3 };

constexpr int FAST_TILE = 2048;  // samples per tile of the fast path; the event-start bitmap is tiled the same way

// A list of reads to be processed by the sequential-order kernels.
//   list  : read indices
//   sbase : base of each entry in the per-sample scratch arrays
//   count : device pointer to the list length
struct WorkList {
    const uint32_t* list;
    const uint64_t* sbase;
    const uint32_t* count;
};

// Device workspace owned by a context.
struct Scratch {
    // sequential-order scratch (gen_cap samples)
    double* Sinc; double* Qinc; float* t1; float* t2;
    uint64_t gen_cap;
    // event-start bitmap over the flat sample span: bit p set <=> an event starts at flat sample p
    uint32_t* bitmap; uint64_t bitmap_words;
    // per tile
    uint32_t* tile_cnt; uint64_t* tile_base;
    uint32_t* tile_read0;            // first read that can intersect a 2048-sample tile (emit_events_kernel); the
                                     // walker sets the top bit of tiles that hold LOW samples (pA <= 0 or barely above)
    // chunk walker (walk.cu)
    uint32_t* wk_cnt;                // [max_reads] interior chunks per read
    uint64_t* wk_ibase;              // [max_reads+1] exclusive scan of wk_cnt
    int* wk_begin; int* wk_end;      // [wk_slots*8] detector state after the warm-up / after the last step of a chunk
    uint64_t wk_slots;
    int* jobs;                       // [job_cap * 4] lives of the long detector to replay: {read, chunk, l_start, end}
    uint32_t* job_count;             // device scalar
    uint32_t job_cap;
    // development / test parameters (sgpu_set_param): 0 = automatic
    uint32_t tune_chunk_len, tune_warmup;
    float tune_thr_long;             // the long detector's threshold (9.0)
    uint32_t tune_stat_cta_min;      // reads of at least this many samples get a CTA in the moments kernels (stat.cu)
    uint32_t max_tiles;
    // per read
    uint32_t* wit_min; uint32_t* wit_max;  // exact-sum witness: min nonzero |pA| / max |pA| bit patterns
    uint32_t* seq_list;     // compacted list of reads routed to the sequential-order kernels
    uint64_t* seq_sbase;
    uint32_t* seq_count;    // device scalar
    unsigned long long* cursor;  // device scalar: used scratch samples
    // scan
    unsigned long long* scan_status; uint32_t* scan_ticket;
    // device-side status / counters
    int* status; unsigned long long* counters;  // [0]=n_events [1]=n_seq [2]=n_fixups
};

// generic.cu (sequential-order kernels)
int launch_generic_detect(const DevBatch& b, const WorkList& wl, Scratch& sc, int sm_count, cudaStream_t st);
int launch_generic_emit(const DevBatch& b, const WorkList& wl, Scratch& sc, const uint64_t* ev_off, uint64_t ev_cap,
                        uint32_t* ev_start, float* ev_mean, float* ev_stdv, int sm_count, cudaStream_t st);
int launch_scan_u32(const uint32_t* cnt, uint32_t n, uint64_t* off, uint64_t* total_out, Scratch& sc, cudaStream_t st);
uint32_t scan_tiles_for(uint32_t n);

// fast.cu (witness evaluation, event ranks) and emit.cu (event table)
uint32_t fast_tiles_for(uint64_t span);
int launch_init_reads(const DevBatch& b, Scratch& sc, uint32_t* seq_flag, uint32_t* fixups, int sm_count,
                      cudaStream_t st);
// walk.cu (chunk walker: pA, t-statistics, peak detector -> event-start bitmap)
int launch_walk(const DevBatch& b, Scratch& sc, float* pa_out, uint32_t* seq_flag, int sm_count, cudaStream_t st);
int launch_long_jobs(const DevBatch& b, Scratch& sc, uint32_t* seq_flag, int sm_count, cudaStream_t st);
int launch_verify_chunks(const DevBatch& b, Scratch& sc, uint32_t* seq_flag, uint32_t* fixups, int sm_count, cudaStream_t st);
uint64_t walk_state_slots(uint64_t max_samples, uint32_t max_reads);
uint32_t walk_chunk_len(uint64_t span, uint32_t n_reads, int rna, int sm_count, uint32_t forced);
uint32_t walk_warmup(int rna, uint32_t forced);
uint32_t walk_job_capacity(uint64_t max_samples);
int launch_build_seq_list(const DevBatch& b, Scratch& sc, uint32_t* seq_flag, int force_all, int sm_count,
                          cudaStream_t st);
int launch_rank_events(const DevBatch& b, Scratch& sc, uint64_t* ev_off, int sm_count, cudaStream_t st);
int launch_fast_emit(const DevBatch& b, Scratch& sc, uint64_t ev_cap, uint32_t* ev_start, float* ev_mean,
                     float* ev_stdv, const uint32_t* fixups, int sm_count, cudaStream_t st);
int launch_pa(const DevBatch& b, float* pa, int sm_count, cudaStream_t st);

// svbzd.cu (svb-zd signal streams -> int16 samples in HBM)
struct SvbBatch {
    const uint8_t*  bytes;     // the streams of the batch back to back, each starting on a 16-byte boundary;
    uint64_t        n_bytes;   // the allocation extends at least 8 bytes past n_bytes
    const uint64_t* comp_off;  // [n_reads] byte offset of every stream
    const uint32_t* comp_len;  // [n_reads] stream length in bytes (header + keys + data)
    const uint64_t* read_off;  // [n_reads+1] where the samples go (multiples of 8)
    const uint32_t* read_len;  // [n_reads] = the count in every stream's header
    uint32_t n_reads;
    uint64_t n_blocks;         // sum of svbzd_blocks_of(read_len)
};
struct SvbScratch {
    uint32_t* cnt; uint64_t* base;            // [max_reads], [max_reads+1]
    uint32_t* blk_bytes; uint64_t* blk_gpos;  // [max_blocks], [max_blocks+1]
    uint32_t* blk_sum; uint64_t* blk_vpos;
    uint32_t* lane_sum;                       // [max_blocks*32]
    uint32_t* blk_read;                       // [max_blocks] read of every block
    uint64_t max_blocks;
};
uint64_t svbzd_max_blocks(uint64_t max_samples, uint32_t max_reads);
uint32_t svbzd_blocks_of(uint64_t n);
int launch_svbzd_decode(const SvbBatch& s, SvbScratch& w, Scratch& sc, int16_t* samples, int sm_count, cudaStream_t st);

// stat.cu
constexpr uint32_t STAT_CTA_MIN_DEFAULT = 131072u;
int launch_stat_moments(const DevBatch& b, float* stat6, uint32_t cta_min, int sm_count, cudaStream_t st);
int launch_stat_median(const DevBatch& b, float* stat6, uint32_t cta_min, int sm_count, cudaStream_t st);
int launch_jnn_moments(const DevBatch& b, float* moments2, uint32_t cta_min, int sm_count, cudaStream_t st);

// jnn.cu (`sigtk jnn`: band from the clamped signal's mean / stdv, counter machine per read -> (start, end) pairs)
uint64_t jnn_seg_capacity(uint64_t max_samples, uint32_t max_reads);
int launch_jnn(const DevBatch& b, const float* moments, uint32_t* seg_cnt, int32_t* seg, int sm_count, cudaStream_t st);


// prefix.cu (`sigtk prefix`: adaptor by jnnv2 on the rolling mean, poly-A by jnn_core on pA, statistics of both)
int launch_prefix(const DevBatch& b, int rna004, int32_t* pos4, float* st6, int sm_count, cudaStream_t st);

// ent.cu (`sigtk ent`: entropies of the raw samples, their zig-zag deltas and the deltas' byte planes)
uint64_t ent_overflow_words(int sm_count);
int launch_ent(const DevBatch& b, uint32_t* overflow, double* out3, int sm_count, cudaStream_t st);

}  // namespace sgpu
