// rsqrt_err.cu -- measures the maximum relative error of rsqrt.approx.ftz.f32 (MUFU.RSQ) on this GPU over EVERY
// positive normal float, against 1/sqrt in double. The walker's t-statistic tail (walk_core.cuh, tail()) relies on
// the bound PTX states (2^-22.4); oracle/proofs/tstat_tail_check.c validates the guard under that bound.
// Build: nvcc -arch=sm_100a -O3 rsqrt_err.cu -o rsqrt_err
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>

__global__ void k(unsigned long long* worst_bits) {
    double worst = 0.0;
    const uint64_t n = 0x7f800000ull - 0x00800000ull;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float a = __uint_as_float((uint32_t)(i + 0x00800000ull));
        float y;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a));
        const double e = fabs((double)y * sqrt((double)a) - 1.0);
        worst = e > worst ? e : worst;
    }
    atomicMax(worst_bits, (unsigned long long)__double_as_longlong(worst));
}

int main() {
    unsigned long long* d;
    cudaMalloc(&d, 8);
    cudaMemset(d, 0, 8);
    k<<<148 * 16, 256>>>(d);
    unsigned long long h = 0;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    double w;
    memcpy(&w, &h, 8);
    printf("{\"what\": \"max relative error of rsqrt.approx.ftz.f32 over all positive normal floats\", \"value\": %.6e, \"log2\": %.3f, \"bound_used\": \"2^-22.4 = 1.81e-7\", \"ok\": %s}\n",
           w, log2(w), w <= 1.81e-7 ? "true" : "false");
    return w <= 1.81e-7 ? 0 : 1;
}
