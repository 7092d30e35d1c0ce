// kernels.cuh -- internal interfaces between the kernel translation units and capi.cu
#pragma once
#include "common.cuh"

namespace sgpu {

// device-side status codes (mapped to SGPU_E_* by capi.cu)
enum { SGPU_DEV_OK = 0, SGPU_DEV_E_EVCAP = 1, SGPU_DEV_E_SCRATCH = 2 };

// A list of reads to be processed by the sequential-order kernels.
//   list  : read indices (nullptr = identity 0..n-1)
//   sbase : [n+1] base of each entry in the per-sample scratch arrays; sbase[n] = used scratch span
//   count : device pointer to the list length (nullptr = n_fixed)
struct WorkList {
    const uint32_t* list;
    const uint64_t* sbase;
    const uint32_t* count;
    uint32_t n_fixed;  // host-known upper bound on the length (sizes the grid)
};

// Device workspace owned by a context.
struct Scratch {
    // sequential-order scratch (gen_cap samples)
    double* Sinc; double* Qinc; float* t1; float* t2;
    uint64_t gen_cap;
    // peak bitmap over the flat sample span: bit p set <=> an event starts at flat sample p
    uint32_t* bitmap; uint64_t bitmap_words;
    // per-read
    uint32_t* ev_cnt;       // events per read
    uint32_t* seq_flag;     // 1 = read needs the sequential-order kernels
    uint32_t* fixups;       // detector chunk fix-ups per read
    uint32_t* seq_list;     // compacted list of flagged reads
    uint64_t* seq_sbase;    // [max_reads+1]
    uint32_t* seq_count;    // device scalar
    // scan
    unsigned long long* scan_status; uint32_t* scan_ticket;
    // device-side status / counters
    int* status; unsigned long long* counters;  // [0]=n_events [1]=n_seq [2]=n_fixups
};

// generic.cu
int launch_generic_detect(const DevBatch& b, const WorkList& wl, uint64_t scratch_span_hint, Scratch& sc,
                          int clear_first, int sm_count, cudaStream_t st);
int launch_count_scan(const DevBatch& b, Scratch& sc, uint64_t* ev_off, uint64_t* total_out, int sm_count,
                      cudaStream_t st);
int launch_generic_emit(const DevBatch& b, const WorkList& wl, Scratch& sc, const uint64_t* ev_off, uint64_t ev_cap,
                        uint32_t* ev_start, float* ev_mean, float* ev_stdv, int* status, int sm_count,
                        cudaStream_t st);
int launch_pa(const DevBatch& b, float* pa, int sm_count, cudaStream_t st);
uint32_t scan_tiles_for(uint32_t n_reads);

// stat.cu
int launch_stat(const DevBatch& b, float* stat6, int sm_count, cudaStream_t st);

}  // namespace sgpu
