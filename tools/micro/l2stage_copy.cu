// Microbenchmark: the MEMORY side of an "L2-staged" K3 pass 1 at 2^24 with no arithmetic.  Instead of landing a
// [4096 x 2] column tile in shared memory (16-byte rows from HBM), a CTA owns a block of C columns (8 C-byte rows)
// and runs the 4096-point column transform as TWO register-resident radix-64 phases (n1 = 64 a + b):
//   phase 1: thread (b, n2) loads x[64 a + b][n2], a = 0..63  (rows of C*8 contiguous bytes), [FFT over a], stores
//            Y[ka][b][n2] into the CTA's PRIVATE scratch in global memory (4096 * C * 8 bytes, reused for every block:
//            it stays in L2)
//   phase 2: thread (ka, n2) loads Y[ka][b][n2], b = 0..63 from the scratch, [FFT over b, twiddle], stores
//            mid[block][ka + 64 kb][n2]: contiguous per block
// This kernel only MOVES the data that way.  It answers: can the strided side be read in 64/128-byte rows at copy
// speed when the exchange goes through L2 instead of shared memory?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2stage_copy l2stage_copy.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
constexpr int N1 = 4096, N2 = 4096;

__device__ __forceinline__ uint64_t policy_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float2 ld_hint(const float2 *p, uint64_t pol) {
    float2 r;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(r.x), "=f"(r.y) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void st_hint(float2 *p, float2 v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}

template <int C, int THREADS, int MINB, bool HINT>
__global__ void __launch_bounds__(THREADS, MINB) l2stage(const float2 *__restrict__ in, float2 *__restrict__ mid,
                                                         float2 *__restrict__ scratch, int n_blocks) {
    float2 *sc = scratch + (size_t)blockIdx.x * N1 * C;
    const uint64_t pf = policy_first(), pl = policy_last();
    constexpr int TASKS = 64 * C;                 // (b, n2) resp. (ka, n2) pairs per phase
    for (int j = blockIdx.x; j < n_blocks; j += gridDim.x) {
        for (int task = threadIdx.x; task < TASKS; task += THREADS) {
            const int b = task / C, n2 = task % C;
            const float2 *src = in + (size_t)b * N2 + (size_t)j * C + n2;
            float2 v[64];
#pragma unroll
            for (int a = 0; a < 64; ++a) v[a] = HINT ? ld_hint(src + (size_t)a * 64 * N2, pf) : __ldg(src + (size_t)a * 64 * N2);
            float2 *dst = sc + b * C + n2;
#pragma unroll
            for (int a = 0; a < 64; ++a) { if (HINT) st_hint(dst + a * 64 * C, v[a], pl); else dst[a * 64 * C] = v[a]; }
        }
        __syncthreads();
        for (int task = threadIdx.x; task < TASKS; task += THREADS) {
            const int ka = task / C, n2 = task % C;
            const float2 *src = sc + ka * 64 * C + n2;
            float2 v[64];
#pragma unroll
            for (int b = 0; b < 64; ++b) v[b] = HINT ? ld_hint(src + b * C, pl) : __ldcg(src + b * C);
            float2 *dst = mid + (size_t)j * N1 * C + ka * C + n2;
#pragma unroll
            for (int b = 0; b < 64; ++b) { if (HINT) st_hint(dst + b * 64 * C, v[b], pf); else dst[b * 64 * C] = v[b]; }
        }
        __syncthreads();
    }
}

template <int C, int THREADS, int MINB, bool HINT> static int run(const float2 *in, float2 *mid, float2 *scratch) {
    const int blocks = N2 / C, grid = 148 * MINB < blocks ? 148 * MINB : blocks;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) l2stage<C, THREADS, MINB, HINT><<<grid, THREADS>>>(in, mid, scratch, blocks);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) l2stage<C, THREADS, MINB, HINT><<<grid, THREADS>>>(in, mid, scratch, blocks);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 20;
    printf("C=%2d (%3d-byte rows), %3d threads x %d CTA/SM, scratch in flight %5.1f MB, hints %d: %.4f ms = %.0f GB/s read + written\n",
           C, C * 8, THREADS, MINB, (double)grid * N1 * C * 8 / 1e6, (int)HINT, ms, 2.0 * N1 * N2 * 8 / ms / 1e6);
    return 0;
}

int main() {
    float2 *in, *mid, *scratch;
    CK(cudaMalloc(&in, (size_t)N1 * N2 * 8));
    CK(cudaMalloc(&mid, (size_t)N1 * N2 * 8));
    CK(cudaMalloc(&scratch, (size_t)148 * 4 * N1 * 16 * 8));
    CK(cudaMemset(in, 0, (size_t)N1 * N2 * 8));
    if (run<8, 128, 1, false>(in, mid, scratch)) return 1;
    if (run<8, 128, 2, false>(in, mid, scratch)) return 1;
    if (run<8, 128, 4, false>(in, mid, scratch)) return 1;
    if (run<8, 256, 1, false>(in, mid, scratch)) return 1;
    if (run<8, 256, 2, false>(in, mid, scratch)) return 1;
    if (run<16, 128, 2, false>(in, mid, scratch)) return 1;
    if (run<16, 256, 1, false>(in, mid, scratch)) return 1;
    if (run<16, 256, 2, false>(in, mid, scratch)) return 1;
    if (run<4, 128, 4, false>(in, mid, scratch)) return 1;
    if (run<8, 128, 2, true>(in, mid, scratch)) return 1;
    if (run<8, 128, 4, true>(in, mid, scratch)) return 1;
    if (run<8, 256, 2, true>(in, mid, scratch)) return 1;
    if (run<16, 128, 2, true>(in, mid, scratch)) return 1;
    if (run<16, 256, 2, true>(in, mid, scratch)) return 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) cudaMemcpyAsync(mid, in, (size_t)N1 * N2 * 8, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) cudaMemcpyAsync(mid, in, (size_t)N1 * N2 * 8, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("cudaMemcpy D2D of the same 134 MB: %.4f ms\n", ms / 20);
    return 0;
}
