// Microbenchmark: throughput of packed FADD2/FFMA2 vs scalar FADD/FFMA on B200.  nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float addv(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fmav(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

template <int MODE> __global__ void k(float *out, int iters, float seed) {
    float s[16]; u64 p[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = seed + i + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = ((u64)__float_as_uint(s[2 * i]) << 32) | __float_as_uint(s[2 * i + 1]);
    const float c = seed * 0.5f; const u64 c2 = ((u64)__float_as_uint(c) << 32) | __float_as_uint(c);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (MODE == 0) { _Pragma("unroll") for (int i = 0; i < 16; ++i) s[i] = addv(s[i], c); }
            if (MODE == 1) { _Pragma("unroll") for (int i = 0; i < 8; ++i) p[i] = add2(p[i], c2); }
            if (MODE == 2) { _Pragma("unroll") for (int i = 0; i < 16; ++i) s[i] = fmav(s[i], c, c); }
            if (MODE == 3) { _Pragma("unroll") for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], c2, c2); }
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += s[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) r += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> float run(float *d, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 4, 256>>>(d, iters, 1.0f);
    cudaEventRecord(e0); k<MODE><<<148 * 4, 256>>>(d, iters, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    float *d; cudaMalloc(&d, 148 * 4 * 256 * 4);
    const int iters = 20000;
    const double flops_lane = 32.0 * iters * 4 * 16;   // per SM: 32 warps x 16 f32 ops per inner pass
    float t0 = run<0>(d, iters), t1 = run<1>(d, iters), t2 = run<2>(d, iters), t3 = run<3>(d, iters);
    printf("FADD  : %.3f ms  %.1f f32 adds/clk/SM\n", t0, flops_lane * 32 / (t0 * 1e-3 * 1.965e9));
    printf("FADD2 : %.3f ms  %.1f f32 adds/clk/SM\n", t1, flops_lane * 32 / (t1 * 1e-3 * 1.965e9));
    printf("FFMA  : %.3f ms  %.1f f32 fmas/clk/SM\n", t2, flops_lane * 32 / (t2 * 1e-3 * 1.965e9));
    printf("FFMA2 : %.3f ms  %.1f f32 fmas/clk/SM\n", t3, flops_lane * 32 / (t3 * 1e-3 * 1.965e9));
    return 0;
}
