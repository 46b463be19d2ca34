// Microbenchmark: do SHFL and LDS share the LSU data pipe on B200?  Build: nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>   // 0: LDS.128 only, 1: SHFL only, 2: both interleaved (same counts as 0 + 1)
__global__ void k(float *out, int iters) {
    __shared__ float4 s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    const int l = threadIdx.x;
    float4 acc = make_float4(0, 0, 0, 0);
    float v0 = l, v1 = l + 1, v2 = l + 2, v3 = l + 3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0 || MODE == 2) {
                float4 t = s[(l + 32 * u + it) & 1023];
                acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
            }
            if (MODE == 1 || MODE == 2) {
                v0 = __shfl_xor_sync(0xffffffffu, v0, 1 + (u & 3));
                v1 = __shfl_xor_sync(0xffffffffu, v1, 2);
                v2 = __shfl_xor_sync(0xffffffffu, v2, 4);
                v3 = __shfl_xor_sync(0xffffffffu, v3, 8);
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w + v0 + v1 + v2 + v3;
}

template <int MODE> float run(float *d, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 4, 256>>>(d, iters);
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 256>>>(d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    float *d; cudaMalloc(&d, 148 * 4 * 256 * 4);
    const int iters = 20000;
    float a = run<0>(d, iters), b = run<1>(d, iters), c = run<2>(d, iters);
    // per SM: 32 warps, each iteration 8 LDS.128 (4 wavefronts each) and 32 SHFL
    double wf = 32.0 * iters * 8 * 4, sh = 32.0 * iters * 32;
    printf("LDS.128 only : %.3f ms  -> %.2f wavefronts/clk/SM @1.965GHz\n", a, wf / (a * 1e-3 * 1.965e9));
    printf("SHFL only    : %.3f ms  -> %.2f shfl/clk/SM\n", b, sh / (b * 1e-3 * 1.965e9));
    printf("both         : %.3f ms  (sum %.3f, max %.3f)\n", c, a + b, a > b ? a : b);
    return 0;
}
