// pipes.cu -- development aid: issue cost (cycles per warp-instruction per SM sub-partition) of the instruction
// classes the chunk walker is made of, alone and mixed, on the GPU it runs on. Build: nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096
#define REP 8   // independent chains per thread

template <int MODE>
__global__ void __launch_bounds__(1024) k(double* out, long long* cyc, double seed, int iters) {
    double d[REP]; float f[REP]; uint32_t u[REP];
#pragma unroll
    for (int i = 0; i < REP; i++) { d[i] = seed + i + threadIdx.x; f[i] = (float)seed + i; u[i] = threadIdx.x * 7 + i; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < REP; i++) {
            if (MODE == 0) d[i] = __fma_rn(d[i], 1.0000001, 0.5);                       // DFMA
            if (MODE == 1) { f[i] = __double2float_rn(d[i]); d[i] = __longlong_as_double(__double_as_longlong(d[i]) + __float_as_int(f[i])); }  // F2F.F32.F64 + IADD
            if (MODE == 2) u[i] = (u[i] & 0x55555555u) ^ (u[i] >> 3) ;                    // ALU (LOP3/SHF)
            if (MODE == 3) f[i] = __fmaf_rn(f[i], 1.0001f, 0.5f);                        // FFMA
            if (MODE == 4) { asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(f[i])); }  // MUFU
            if (MODE == 5) { d[i] = __fma_rn(d[i], 1.0000001, 0.5); u[i] = (u[i] & 0x55555555u) ^ (u[i] >> 3); }  // DFMA + 2 ALU
            if (MODE == 6) { f[i] = __double2float_rn(d[i]); d[i] = __fma_rn(d[i], 1.0000001, (double)0.5); u[i] ^= __float_as_uint(f[i]); }  // F2F + DFMA + LOP
            if (MODE == 7) { asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(f[i])); d[i] = __fma_rn(d[i], 1.0000001, 0.5); }  // MUFU + DFMA
            if (MODE == 8) { f[i] = __fmaf_rn(f[i], 1.0001f, 0.5f); u[i] = (u[i] & 0x55555555u) ^ (u[i] >> 3); }  // FFMA + 2 ALU
            if (MODE == 9) { unsigned long long w = (unsigned long long)u[i] * 0x20000000ull + 0x3800000000000000ull; u[i] = (uint32_t)(w >> 32) ^ (uint32_t)w; }  // IMAD.WIDE + LOP
            if (MODE == 10) { f[i] = __double2float_rn(d[i]); d[i] = __longlong_as_double(__double_as_longlong(d[i]) + 1); u[i] = (u[i] & 0x55555555u) ^ (u[i] >> 3) ^ __float_as_uint(f[i]); }  // F2F + 4 ALU
            if (MODE == 11) { float g = f[i], h = f[i] + 1.0f; asm volatile("{.reg .b64 a; mov.b64 a, {%0,%1}; fma.rn.f32x2 a, a, a, a; mov.b64 {%0,%1}, a;}" : "+f"(g), "+f"(h)); f[i] = g + h; }  // FFMA2 + FADD
            if (MODE == 12) { d[i] = (double)f[i]; f[i] = __int_as_float(__double2hiint(d[i])); }  // F2F.F64.F32
            if (MODE == 14) u[i] = u[i] * 3u + 7u;                                         // IMAD (32-bit)
            if (MODE == 15) { const uint32_t b = __float_as_uint(f[i]); const double w = __hiloint2double((int)((b >> 3) + 0x38000000u), (int)(b << 29)); d[i] = __dadd_rn(d[i], w); }  // 2-op widen + DADD
            if (MODE == 16) { const double w = __longlong_as_double((long long)((unsigned long long)__float_as_uint(f[i]) * 0x20000000ull + 0x3800000000000000ull)); d[i] = __dadd_rn(d[i], w); }  // IMAD.WIDE widen + DADD
            if (MODE == 17) { unsigned long long w = __double_as_longlong(d[i]); w = (unsigned long long)u[i] * 0x20000000ull + w; d[i] = __longlong_as_double(w); }  // IMAD.WIDE chain
            if (MODE == 18) { d[i] = __dadd_rn(d[i], (double)f[i]); }  // F2F.F64.F32 + DADD
            if (MODE == 13) { f[i] = fmaxf(f[i], (float)i) ; u[i] += (f[i] > 3.0f); }  // FMNMX + FSETP+...
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < REP; i++) s += d[i] + f[i] + u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_rep, int blocks_per_sm) {
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 16 * 128 * 8); cudaMalloc(&cyc, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, 128 * blocks_per_sm>>>(out, cyc, 1.5, 64);
    cudaEventRecord(e0);
    k<MODE><<<148, 128 * blocks_per_sm>>>(out, cyc, 1.5, ITER);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    // per SMSP: blocks_per_sm warps (each block = 4 warps, one per SMSP); clock from the in-kernel cycle count of block 0
    const double per_thread = (double)c / ((double)ITER * REP);
    const double ghz = 1.965;
    const double per_smsp = (double)ms * 1e-3 * ghz * 1e9 / ((double)ITER * REP * blocks_per_sm);
    printf("%-22s warps/SMSP=%d  block0: %6.2f cyc/rep/warp | whole GPU: %6.2f cyc/rep/SMSP  (%d instr/rep -> %.2f cyc/instr)\n", name,
           blocks_per_sm, per_thread, per_smsp, instr_per_rep, per_smsp / instr_per_rep);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    const int w = 8;
    run<0>("DFMA", 1, w); run<1>("F2F.F32.F64+IADD", 2, w); run<2>("ALU x2", 2, w); run<3>("FFMA", 1, w); run<4>("MUFU.RSQ", 1, w);
    run<5>("DFMA + 2 ALU", 3, w); run<6>("F2F + DFMA + LOP", 3, w); run<7>("MUFU + DFMA", 2, w); run<8>("FFMA + 2 ALU", 3, w);
    run<9>("IMAD.WIDE + LOP", 2, w); run<10>("F2F + ~5 ALU", 6, w); run<11>("FFMA2 + FADD", 2, w); run<12>("F2F.F64.F32 + MOV", 2, w);
    run<13>("FMNMX+FSETP+..", 3, w);
    run<14>("IMAD32", 1, w); run<15>("widen 2-op + DADD", 3, w); run<16>("widen IMAD.WIDE + DADD", 2, w); run<17>("IMAD.WIDE chain", 1, w); run<18>("F2F.F64.F32 + DADD", 2, w);
    run<0>("DFMA", 1, 1); run<1>("F2F.F32.F64+IADD", 2, 1); run<0>("DFMA", 1, 4); run<1>("F2F.F32.F64+IADD", 2, 4);
    return 0;
}
