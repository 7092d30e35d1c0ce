// host_walk.cpp -- TEST INFRASTRUCTURE ONLY (a checker, never part of the product library).
//
// Runs the chunk walker of sigtk_b200/csrc/walk_core.cuh (the very code the CUDA kernel walk_chunks_kernel
// executes per thread; the header is __host__ __device__ clean) on the CPU, chunk after chunk, so that the
// chunking, register-ring, warm-up and boundary-state logic can be compared with the oracle without a GPU
// (tests/test_host_walk.py). Build: g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC (see tests/_hostwalk.py).
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../sigtk_b200/csrc/walk_core.cuh"

using namespace sgpu::walk;

namespace {
struct HostIo {
    const int16_t* sp;
    float* pa;
    uint32_t* bm;
    Canon* begin;
    Canon* end;
    int* rmin;
    int* rmax;
    void load8(int t, int (&v)[4]) const { memcpy(v, sp + t, 16); }
    bool want_pa() const { return pa != nullptr; }
    int warp_max(int v) const { return v; }
    void prefetch(int) const {}
    void store_pa8(int t, const float* x) const { memcpy(pa + t, x, 32); }
    void store_pa1(int t, float x) const { pa[t] = x; }
    void peak(int pos) const { bm[pos >> 5] |= 1u << (pos & 31); }
    void peaks32(int ub, uint32_t mk) const {
        for (int k = 0; k < 32; k++)
            if (mk >> k & 1u) { if (ub + k < 0) abort(); peak(ub + k); }
    }

    void put_begin(const Canon& c) const { *begin = c; }
    void put_end(const Canon& c) const { *end = c; }
    void witness(int lo, int hi, int) const { if (lo < *rmin) *rmin = lo; if (hi > *rmax) *rmax = hi; }
    void low_samples(int, uint32_t, uint32_t) const {}
    // a life of the long detector that may emit (walk_core.cuh): replayed after all chunks have been walked
    struct Job { int k, l_start, end; };
    std::vector<Job>* jobs;
    int k;
    void job(int l_start, int end) const { jobs->push_back(Job{k, l_start, end}); }
};

// the memory side of long_job(): the END records of the read's chunks
struct HostJobIo {
    const int16_t* sp;
    uint32_t* bm;
    const std::vector<Canon>* end;
    void load8(int t, int (&v)[4]) const { memcpy(v, sp + t, 16); }
    void peak(int pos) const { bm[pos >> 5] |= 1u << (pos & 31); }
    int end_lstart(int kk) const { return (*end)[kk].v[6]; }
    void end_short(int kk, float* pv, int* ps) const { *pv = bits_f((uint32_t)(*end)[kk].v[0]); *ps = (*end)[kk].v[1]; }
};

template <int RNA>
int run(const int16_t* raw_padded, int n, float off, float unit, int L, int W, int sh, uint32_t* bitmap, float* pa,
        int* rminmax, float thr_long, int* n_jobs) {
    const uint32_t nch = n_chunks((uint32_t)n, (uint32_t)L);
    if (nch == 0) return 0;
    std::vector<Canon> begin(nch), end(nch);
    int rmin = 32767, rmax = -32768;
    std::vector<HostIo::Job> jobs;
    for (uint32_t k = 0; k < nch; k++) {
        HostIo io{raw_padded, pa, bitmap, &begin[k], &end[k], &rmin, &rmax, &jobs, (int)k};
        if (k == 0) walk_edge<RNA>(io, n, off, unit, sh, L, W, 0, thr_long);
        else if (k == nch - 1) walk_edge<RNA>(io, n, off, unit, sh, L, W, 1, thr_long);
        else walk_interior<RNA>(io, n, off, unit, sh, L, W, (int)k, thr_long);
    }
    // second pass: the lives of the long detector that may emit, with the reference's own operations
    HostJobIo jo{raw_padded, bitmap, &end};
    for (const HostIo::Job& j : jobs) long_job<RNA>(jo, n, sh, off, unit, L, j.k, j.l_start, j.end, thr_long);
    *n_jobs = (int)jobs.size();
    int mism = 0;
    for (uint32_t k = 1; k < nch; k++) {
        bool same = true;
        for (int q = 0; q < 6; q++) same = same && begin[k].v[q] == end[k - 1].v[q];
        if (!same) mism++;
    }
    rminmax[0] = rmin;
    rminmax[1] = rmax;
    return mism;
}
}  // namespace

// raw_padded: the read's samples followed by enough padding for 16-byte loads up to the 8-aligned length.
// bitmap: ((n + sh + 31) / 32 + 1) zeroed words; bit (i + sh) set <=> an event starts at sample i.
// Returns the number of chunk boundaries whose warm-up state differed from the predecessor's end state.
// thr_long: the long detector's threshold (9.0 in the reference; tests lower it so that the long detector emits).
// n_jobs: number of lives of the long detector that were replayed.
extern "C" int host_walk_read(const int16_t* raw_padded, int n, float off, float unit, int rna, int L, int W, int sh,
                              uint32_t* bitmap, float* pa, int* rminmax, float thr_long, int* n_jobs) {
    return rna ? run<1>(raw_padded, n, off, unit, L, W, sh, bitmap, pa, rminmax, thr_long, n_jobs)
               : run<0>(raw_padded, n, off, unit, L, W, sh, bitmap, pa, rminmax, thr_long, n_jobs);
}
