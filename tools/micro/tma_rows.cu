// Microbenchmark: how fast can one B200 stream a [4096 x 4096] complex-f32 matrix through shared memory in COLUMN
// tiles (all rows, W complex wide), as pass 1 of the four-step FFT must -- with 2-D TMA boxes of W*8-byte rows
// versus plain LDG.128, and the same for the row-tile STORE of pass 2.  No arithmetic: this isolates the memory side.
// Build: nvcc -arch=sm_100a -O3 -o tma_rows tma_rows.cu -lcuda
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
__device__ __forceinline__ void tma_st(const CUtensorMap *m, int x, int y, const void *src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(m), "r"(x), "r"(y), "r"(su32(src)) : "memory");
}

// ---- TMA column-tile load: tile = ROWS rows x W complex; two buffers per CTA, loads always one tile ahead
template <int W, int ROWS>
__global__ void __launch_bounds__(128) tma_load_cols(const __grid_constant__ CUtensorMap map, int n_tiles, int tiles_per_col, float *sink) {
    extern __shared__ __align__(128) unsigned char sm[];
    constexpr int TB = ROWS * W * 8;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm + 2 * TB);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1); mbar_init(bar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    float acc = 0.f;
    if (threadIdx.x == 0) {
        auto issue = [&](int t, int b) {
            mbar_expect(bar + b, TB);
            const int col = t / tiles_per_col, part = t % tiles_per_col;
            for (int r0 = 0; r0 < ROWS; r0 += 256) tma_ld(sm + b * TB + r0 * W * 8, &map, 2 * W * col, part * ROWS + r0, bar + b);
        };
        int t = blockIdx.x, b = 0;
        uint32_t ph[2] = {0, 0};
        if (t < n_tiles) issue(t, 0);
        for (; t < n_tiles; t += gridDim.x) {
            if (t + gridDim.x < n_tiles) issue(t + gridDim.x, b ^ 1);
            mbar_wait(bar + b, ph[b]);
            ph[b] ^= 1;
            acc += reinterpret_cast<float *>(sm + b * TB)[0];
            b ^= 1;
        }
    }
    if (acc == 123.456f) sink[0] = acc;
}

// ---- LDG.128 column-tile load: each thread reads 16 bytes of a row (W = 2 complex per row)
__global__ void __launch_bounds__(512) ldg_load_cols(const float4 *x, int n_rows, int row_f4, int n_tiles, float *sink) {
    float acc = 0.f;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const float4 *p = x + t;          // column pair t: float4 index t within a row
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __ldg(p + (size_t)(threadIdx.x + 512 * i) * row_f4);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    if (acc == 123.456f) sink[0] = acc;
}
// same, but a warp reads 8 rows x 4 adjacent column pairs (64 bytes per row): what wider tiles would allow
__global__ void __launch_bounds__(512) ldg_load_cols64(const float4 *x, int n_rows, int row_f4, int n_tiles, float *sink) {
    float acc = 0.f;
    const int sub = threadIdx.x & 3, r = threadIdx.x >> 2;      // 128 rows per sweep
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {      // tile = 4 column pairs x 1024 rows  (64 KB)
        const int col4 = t % (row_f4 / 4), part = t / (row_f4 / 4);
        const float4 *p = x + 4 * col4 + sub + (size_t)part * 1024 * row_f4;
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __ldg(p + (size_t)(r + 128 * i) * row_f4);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    if (acc == 123.456f) sink[0] = acc;
}

// ---- TMA row-tile store (pass 2 pattern): tile = ROWS rows x W complex written to out[row][col0 .. col0+W)
template <int W, int ROWS>
__global__ void __launch_bounds__(128) tma_store_cols(const __grid_constant__ CUtensorMap map, int n_tiles, int tiles_per_col) {
    extern __shared__ __align__(128) unsigned char sm[];
    constexpr int TB = ROWS * W * 8;
    for (int i = threadIdx.x; i < TB / 4; i += blockDim.x) reinterpret_cast<float *>(sm)[i] = (float)i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int col = t / tiles_per_col, part = t % tiles_per_col;
            for (int r0 = 0; r0 < ROWS; r0 += 256) tma_st(&map, 2 * W * col, part * ROWS + r0, sm + r0 * W * 8);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
// ---- STG.128 row-tile store: W = 2 complex (16 bytes) per row per thread / 64 bytes per row per 4 threads
__global__ void __launch_bounds__(512) stg_store_cols(float4 *out, int row_f4, int n_tiles, int wide) {
    const float4 v = make_float4(threadIdx.x, 1.f, 2.f, 3.f);
    if (!wide) {
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
#pragma unroll
            for (int i = 0; i < 8; ++i) out[t + (size_t)(threadIdx.x + 512 * i) * row_f4] = v;
        }
    } else {
        const int sub = threadIdx.x & 3, r = threadIdx.x >> 2;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int col4 = t % (row_f4 / 4), part = t / (row_f4 / 4);
            float4 *p = out + 4 * col4 + sub + (size_t)part * 1024 * row_f4;
#pragma unroll
            for (int i = 0; i < 8; ++i) p[(size_t)(r + 128 * i) * row_f4] = v;
        }
    }
}

typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                              const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_fn enc;
static bool make_map(CUtensorMap *m, void *base, uint64_t row_floats, uint64_t rows, uint32_t box_floats, uint32_t box_rows) {
    cuuint64_t dims[2] = {row_floats, rows};
    cuuint64_t strides[1] = {row_floats * 4};
    cuuint32_t box[2] = {box_floats, box_rows}, es[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <class F> static float timeit(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / 5;
}

template <int W, int ROWS> static void run_tma(float *d, float *d2, float *sink, int per_sm) {
    constexpr int N = 4096;
    CUtensorMap m, m2;
    make_map(&m, d, 2ull * N, N, 2 * W, 256);
    make_map(&m2, d2, 2ull * N, N, 2 * W, 256);
    const int tiles = (N / W) * (N / ROWS), smem = 2 * ROWS * W * 8 + 64;
    cudaFuncSetAttribute(tma_load_cols<W, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(tma_store_cols<W, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float ms = timeit([&] { tma_load_cols<W, ROWS><<<148 * per_sm, 128, smem>>>(m, tiles, N / ROWS, sink); });
    float ms2 = timeit([&] { tma_store_cols<W, ROWS><<<148 * per_sm, 128, smem>>>(m2, tiles, N / ROWS); });
    printf("TMA  rows of %3d B, tile %4d rows (%3d KB), %d CTA/SM: load %.3f ms = %6.0f GB/s   store %.3f ms = %6.0f GB/s   [%s]\n",
           W * 8, ROWS, ROWS * W * 8 / 1024, per_sm, ms, 134.2 / ms, ms2, 134.2 / ms2, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    constexpr int N = 4096;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    enc = (encode_fn)p;
    float *d, *d2, *sink;
    CK(cudaMalloc(&d, (size_t)N * N * 8));
    CK(cudaMalloc(&d2, (size_t)N * N * 8));
    CK(cudaMalloc(&sink, 16));
    CK(cudaMemset(d, 0, (size_t)N * N * 8));
    run_tma<2, 4096>(d, d2, sink, 1);
    run_tma<2, 2048>(d, d2, sink, 2);
    run_tma<2, 1024>(d, d2, sink, 4);
    run_tma<4, 2048>(d, d2, sink, 1);
    run_tma<4, 1024>(d, d2, sink, 2);
    run_tma<8, 1024>(d, d2, sink, 1);
    run_tma<8, 512>(d, d2, sink, 2);
    run_tma<16, 512>(d, d2, sink, 1);
    run_tma<16, 256>(d, d2, sink, 2);
    for (int per_sm = 1; per_sm <= 4; ++per_sm) {
        float ms = timeit([&] { ldg_load_cols<<<148 * per_sm, 512>>>((const float4 *)d, N, N / 2, N / 2, sink); });
        float ms2 = timeit([&] { ldg_load_cols64<<<148 * per_sm, 512>>>((const float4 *)d, N, N / 2, (N / 8) * 4, sink); });
        float ms3 = timeit([&] { stg_store_cols<<<148 * per_sm, 512>>>((float4 *)d2, N / 2, N / 2, 0); });
        float ms4 = timeit([&] { stg_store_cols<<<148 * per_sm, 512>>>((float4 *)d2, N / 2, (N / 8) * 4, 1); });
        printf("LDG.128 %d CTA/SM: 16 B rows %.3f ms = %6.0f GB/s | 64 B rows %.3f ms = %6.0f GB/s   STG.128: 16 B rows %.3f ms = %6.0f GB/s | 64 B rows %.3f ms = %6.0f GB/s  [%s]\n",
               per_sm, ms, 134.2 / ms, ms2, 134.2 / ms2, ms3, 134.2 / ms3, ms4, 134.2 / ms4, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
