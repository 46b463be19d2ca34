// One large transform over R GPUs: n = R L points, rank r holds the CYCLIC slice x[j R + r]
// (j < L; what a multi-channel interleaved capture looks like per channel).  Decimation in time over the ranks:
//
//   F_r[k'] = sum_j x[j R + r] W_L^{j k'}                                   local L-point transform (K3, no traffic)
//   X[k' + L q] = sum_r W_R^{r q} ( W_n^{r k'} F_r[k'] )                     R-point butterfly ACROSS the ranks
//
// The butterfly is the path's one real exchange step, and it is ONE kernel: rank p finishes the bins k' of its block
// [p L/R, (p+1) L/R) and reads F_r[k'] of every other rank straight from that rank's HBM through CUDA-IPC mapped
// pointers (NVLink peer loads, coalesced: consecutive threads, consecutive k'), applies the twiddle, runs the
// radix-R butterfly in registers and writes its R output rows locally -- exchange and arithmetic overlap thread by
// thread, there is no all-to-all buffer and no second pass.  Output on rank p: out[q][i] = X[(p L/R + i) + L q].
#include <math.h>

#include <algorithm>

#include <mutex>
#include <vector>

#include "common.cuh"
#include "fft_batched.cuh"

namespace kopek {


struct DistParams {
    const float2 *spec[8];    // F_r, r < ranks: this rank's own buffer or a peer mapping
    float2 *out;              // [ranks][block]
    const float2 *tw_hi;      // W_n^{4096 h}
    const float2 *tw_lo;      // W_n^{l}, l < 4096
    uint64_t block0;          // first k' of this rank's block
    uint64_t block;           // L / ranks
};

template <int R>
__global__ void __launch_bounds__(256) dist_combine_kernel(const DistParams p) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.block; i += stride) {
        const uint64_t k = p.block0 + i;
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = p.spec[r][k];                 // r != rank: a load over NVLink
#pragma unroll
        for (int r = 1; r < R; ++r) {
            const uint64_t m = (uint64_t)r * k;                          // < n: no reduction needed
            v[r] = cmul(v[r], cmul(__ldg(p.tw_hi + (m >> 12)), __ldg(p.tw_lo + (m & 4095u))));
        }
        Dft<R>::run(v);
#pragma unroll
        for (int q = 0; q < R; ++q) p.out[(uint64_t)q * p.block + i] = v[q];
    }
}

struct DistTables {
    uint64_t n = 0;
    int device = -1;
    float2 *hi = nullptr, *lo = nullptr;
};
static std::mutex g_dist_mu;
static std::vector<DistTables> g_dist_tables;

static kopek_status dist_tables(uint64_t n, int device, DistTables *out) {
    std::lock_guard<std::mutex> lk(g_dist_mu);
    for (const DistTables &t : g_dist_tables)
        if (t.n == n && t.device == device) { *out = t; return KOPEK_OK; }
    DistTables t;
    t.n = n; t.device = device;
    std::vector<float2> hi(n / 4096), lo(4096);
    for (uint64_t h = 0; h < hi.size(); ++h) {
        const double a = -2.0 * M_PI * (double)(h * 4096) / (double)n;
        hi[h] = make_float2((float)cos(a), (float)sin(a));
    }
    for (uint64_t l = 0; l < 4096; ++l) {
        const double a = -2.0 * M_PI * (double)l / (double)n;
        lo[l] = make_float2((float)cos(a), (float)sin(a));
    }
    if (cudaMalloc(&t.hi, hi.size() * sizeof(float2)) != cudaSuccess || cudaMalloc(&t.lo, lo.size() * sizeof(float2)) != cudaSuccess ||
        cudaMemcpy(t.hi, hi.data(), hi.size() * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(t.lo, lo.data(), lo.size() * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(t.hi); cudaFree(t.lo);
        return fail(KOPEK_ERR_NOMEM, "twiddle tables for n=%llu: %s", (unsigned long long)n, cudaGetErrorString(cudaGetLastError()));
    }
    g_dist_tables.push_back(t);
    *out = t;
    return KOPEK_OK;
}

}  // namespace kopek

using namespace kopek;

extern "C" kopek_status kopek_fft_dist_combine(const float *const *d_spectra, uint32_t ranks, uint32_t rank,
                                               uint64_t local_n, float *d_out, int device, void *cuda_stream) {
    if (!d_spectra || !d_out) return fail(KOPEK_ERR_INVALID, "NULL argument");
    if (ranks != 2 && ranks != 4 && ranks != 8) return fail(KOPEK_ERR_UNSUPPORTED, "ranks=%u: 2, 4 or 8", ranks);
    if (rank >= ranks) return fail(KOPEK_ERR_INVALID, "rank %u of %u", rank, ranks);
    if (local_n < 4096 || (local_n & (local_n - 1)) || local_n % ranks)
        return fail(KOPEK_ERR_INVALID, "local_n=%llu must be a power of two >= 4096", (unsigned long long)local_n);
    for (uint32_t r = 0; r < ranks; ++r)
        if (!d_spectra[r] || (uintptr_t)d_spectra[r] % 8) return fail(KOPEK_ERR_INVALID, "spectrum pointer %u is NULL or misaligned", r);
    DeviceGuard g(device);
    if (!g.ok) return fail(KOPEK_ERR_CUDA, "cannot select device %d", device);
    DistTables t;
    kopek_status st = dist_tables((uint64_t)ranks * local_n, device, &t);
    if (st != KOPEK_OK) return st;
    DistParams p{};
    for (uint32_t r = 0; r < ranks; ++r) p.spec[r] = reinterpret_cast<const float2 *>(d_spectra[r]);
    p.out = reinterpret_cast<float2 *>(d_out);
    p.tw_hi = t.hi; p.tw_lo = t.lo;
    p.block = local_n / ranks;
    p.block0 = (uint64_t)rank * p.block;
    int sms = 0;
    KOPEK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const unsigned grid = (unsigned)std::min<uint64_t>((p.block + 255) / 256, (uint64_t)sms * 8);
    cudaStream_t s = (cudaStream_t)cuda_stream;
    if (ranks == 2) dist_combine_kernel<2><<<grid, 256, 0, s>>>(p);
    else if (ranks == 4) dist_combine_kernel<4><<<grid, 256, 0, s>>>(p);
    else dist_combine_kernel<8><<<grid, 256, 0, s>>>(p);
    KOPEK_CUDA(cudaGetLastError());
    return KOPEK_OK;
}
