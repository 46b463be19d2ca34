// Microbenchmark: the MEMORY side of K3 pass 1 at 2^24 with no arithmetic -- persistent CTAs load [4096 x W] column tiles
// with 2-D TMA boxes and store them tile-major into `mid` exactly as k3_pass1 does (W = 2: 16-byte rows in, 64-byte
// segments out, two CTAs per SM; W = 4: 32-byte rows, 128-byte segments, one CTA per SM).  The next tile is requested
// as soon as the current one has been read into registers.  This is the floor any pass-1 transform can reach.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pass1_copy pass1_copy.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
__device__ __forceinline__ uint32_t su32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t ph) {
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(su32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma_ld(void *dst, const CUtensorMap *m, int x, int y, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(su32(dst)), "l"(m), "r"(x), "r"(y), "r"(su32(bar)) : "memory");
}

constexpr int N1 = 4096, N2 = 4096;

template <int W, int THREADS, int PER>     // THREADS * PER = N1 * W
__global__ void __launch_bounds__(THREADS) pass1_copy(const __grid_constant__ CUtensorMap map, float2 *mid, int n_tiles) {
    extern __shared__ __align__(128) unsigned char sm[];
    float2 *buf = reinterpret_cast<float2 *>(sm);
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm + N1 * W * 8);
    constexpr int TB = N1 * W * 8, T = THREADS / W;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if ((int)blockIdx.x < n_tiles) {
            mbar_expect(bar, TB);
            for (int r0 = 0; r0 < N1; r0 += 256) tma_ld(buf + r0 * W, &map, 2 * W * blockIdx.x, r0, bar);
        }
    }
    __syncthreads();
    const int c = threadIdx.x % W, tid = threadIdx.x / W;
    const size_t seg_stride = (size_t)n_tiles * 4 * W;
    uint32_t ph = 0;
    for (int s = blockIdx.x; s < n_tiles; s += gridDim.x) {
        mbar_wait(bar, ph);
        ph ^= 1;
        float2 v[PER];
#pragma unroll
        for (int k = 0; k < PER; ++k) v[k] = buf[(tid + T * k) * W + c];
        __syncthreads();
        const int nxt = s + gridDim.x;
        if (threadIdx.x == 0 && nxt < n_tiles) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect(bar, TB);
            for (int r0 = 0; r0 < N1; r0 += 256) tma_ld(buf + r0 * W, &map, 2 * W * nxt, r0, bar);
        }
        float2 *dst = mid + (size_t)s * 4 * W + (tid & 3) * W + c + (size_t)(tid >> 2) * seg_stride;
#pragma unroll
        for (int k = 0; k < PER; ++k) dst[(size_t)(T / 4) * k * seg_stride] = v[k];
    }
}

typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                              const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int W, int THREADS, int PER> static int run(encode_fn enc, const float2 *in, float2 *mid, int ctas_per_sm) {
    CUtensorMap m;
    cuuint64_t dims[2] = {2ull * N2, (cuuint64_t)N1};
    cuuint64_t strides[1] = {2ull * N2 * sizeof(float)};
    cuuint32_t box[2] = {2 * W, 256}, es[2] = {1, 1};
    if (enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 1;
    const int smem = N1 * W * 8 + 16, tiles = N2 / W;
    CK(cudaFuncSetAttribute(pass1_copy<W, THREADS, PER>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int grid = 148 * ctas_per_sm < tiles ? 148 * ctas_per_sm : tiles;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) pass1_copy<W, THREADS, PER><<<grid, THREADS, smem>>>(m, mid, tiles);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) pass1_copy<W, THREADS, PER><<<grid, THREADS, smem>>>(m, mid, tiles);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 20;
    printf("W=%d (%2d-byte rows in, %3d-byte segments out), %4d threads x %d CTA/SM: %.4f ms = %.0f GB/s read + written\n", W, W * 8,
           W * 32, THREADS, ctas_per_sm, ms, 2.0 * N1 * N2 * 8 / ms / 1e6);
    return 0;
}

int main() {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    encode_fn enc = (encode_fn)p;
    float2 *in, *mid;
    CK(cudaMalloc(&in, (size_t)N1 * N2 * 8));
    CK(cudaMalloc(&mid, (size_t)N1 * N2 * 8));
    CK(cudaMemset(in, 0, (size_t)N1 * N2 * 8));
    if (run<2, 512, 16>(enc, in, mid, 2)) return 1;
    if (run<2, 128, 64>(enc, in, mid, 3)) return 1;
    if (run<4, 256, 64>(enc, in, mid, 1)) return 1;
    if (run<4, 512, 32>(enc, in, mid, 1)) return 1;
    // plain copy of the same bytes as the yardstick
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
