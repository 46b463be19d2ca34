// Instantiations + host launchers of the sub-warp kernels: 256 points (fft_warp256.cuh, eight lanes per frame) and
// 512 points (fft_warp512.cuh, sixteen lanes per frame).
#include <math.h>

#include "k1w_launch.cuh"
#include "fft_warp512.cuh"

namespace kopek {

// ------------------------------------------------------------------ 256-point, eight lanes per frame ----
void k1w256_build_tables(const float *half_window /* 256 or nullptr */, float4 *out) {
    const double tau = 2.0 * M_PI;
    for (int i = 0; i < K256_WIN_F4; ++i)
        out[i] = half_window ? make_float4(half_window[4 * i], half_window[4 * i + 1], half_window[4 * i + 2],
                                           half_window[4 * i + 3])
                             : make_float4(0.5f, 0.5f, 0.5f, 0.5f);
    float4 *tw = out + K256_WIN_F4;
    for (int k1 = 1; k1 < 8; ++k1)
        for (int a = 0; a < 8; ++a) {
            double x = -tau * (double)(2 * a * k1) / 128.0, y = -tau * (double)((2 * a + 1) * k1) / 128.0;
            tw[(k1 - 1) * 8 + a] = make_float4((float)cos(x), (float)sin(x), (float)cos(y), (float)sin(y));
        }
    float4 *pw = tw + K256_TW_F4;
    for (int a = 0; a < 8; ++a) {
        double x = tau * (double)a / 256.0;
        pw[a] = make_float4((float)-sin(x), (float)-cos(x), 0.0f, 0.0f);
    }
}

template <int OUT, bool WIN, bool FULL, bool VEC>
static cudaError_t launch256_v(const K1WParams &p, int grid, cudaStream_t stream, int *occ_out) {
    constexpr int WARPS = 4;
    auto kern = k1w_256_kernel<OUT, WIN, WARPS, 3, FULL, VEC>;
    constexpr size_t smem = sizeof(float4) * WARPS * K256_EX_F4;
    if (occ_out) {
        int nb = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, WARPS * 32, smem);
        *occ_out = nb * WARPS;
        return e;
    }
    kern<<<grid, WARPS * 32, smem, stream>>>(p);
    return cudaGetLastError();
}

// mode: bit 0 = FULL rows, bit 1 = 32-bit loads (odd hop / unaligned base)
template <int OUT, bool WIN>
static cudaError_t launch256(const K1WParams &p, int mode, int grid, cudaStream_t stream, int *occ_out) {
    switch (occ_out ? 0 : mode) {
        case 1: return launch256_v<OUT, WIN, true, true>(p, grid, stream, occ_out);
        case 2: return launch256_v<OUT, WIN, false, false>(p, grid, stream, occ_out);
        case 3: return launch256_v<OUT, WIN, true, false>(p, grid, stream, occ_out);
        default: return launch256_v<OUT, WIN, false, true>(p, grid, stream, occ_out);
    }
}

cudaError_t k1w256_launch(const K1WParams &p, int output, bool win, int mode, int grid, cudaStream_t stream, int *occ_out) {
    switch (output * 2 + (win ? 1 : 0)) {
        case 0: return launch256<OUT_COMPLEX, false>(p, mode, grid, stream, occ_out);
        case 1: return launch256<OUT_COMPLEX, true>(p, mode, grid, stream, occ_out);
        case 2: return launch256<OUT_MAG, false>(p, mode, grid, stream, occ_out);
        case 3: return launch256<OUT_MAG, true>(p, mode, grid, stream, occ_out);
        case 4: return launch256<OUT_POWER, false>(p, mode, grid, stream, occ_out);
        case 5: return launch256<OUT_POWER, true>(p, mode, grid, stream, occ_out);
        case 6: return launch256<OUT_DB, false>(p, mode, grid, stream, occ_out);
        case 7: return launch256<OUT_DB, true>(p, mode, grid, stream, occ_out);
        default: return cudaErrorInvalidValue;
    }
}

// ------------------------------------------------------------------ 512-point, sixteen lanes per frame ----
void k1w512_build_tables(const float *half_window /* 512 or nullptr */, float4 *out4) {
    float2 *out = reinterpret_cast<float2 *>(out4);
    const double tau = 2.0 * M_PI;
    for (int i = 0; i < K512_WIN_F2; ++i)
        out[i] = half_window ? make_float2(half_window[2 * i], half_window[2 * i + 1]) : make_float2(0.5f, 0.5f);
    float2 *tw = out + K512_WIN_F2;
    for (int k1 = 1; k1 < 16; ++k1)
        for (int a = 0; a < 16; ++a) {
            double x = -tau * (double)(a * k1) / 256.0;
            tw[(k1 - 1) * 16 + a] = make_float2((float)cos(x), (float)sin(x));
        }
    float2 *pw = tw + K512_TW_F2;
    for (int a = 0; a < 16; ++a) {
        double x = tau * (double)a / 512.0;
        pw[a] = make_float2((float)-sin(x), (float)-cos(x));
    }
}

template <int OUT, bool WIN, bool FULL, bool VEC>
static cudaError_t launch512_v(const K1WParams &p, int grid, cudaStream_t stream, int *occ_out) {
    constexpr int WARPS = 4;
    auto kern = k1w_512_kernel<OUT, WIN, WARPS, 3, FULL, VEC>;
    constexpr size_t smem = sizeof(float2) * WARPS * K512_EX_F2;
    if (occ_out) {
        int nb = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, WARPS * 32, smem);
        *occ_out = nb * WARPS;
        return e;
    }
    kern<<<grid, WARPS * 32, smem, stream>>>(p);
    return cudaGetLastError();
}

// mode: bit 0 = FULL rows, bit 1 = 32-bit loads (odd hop / unaligned base)
template <int OUT, bool WIN>
static cudaError_t launch512(const K1WParams &p, int mode, int grid, cudaStream_t stream, int *occ_out) {
    switch (occ_out ? 0 : mode) {
        case 1: return launch512_v<OUT, WIN, true, true>(p, grid, stream, occ_out);
        case 2: return launch512_v<OUT, WIN, false, false>(p, grid, stream, occ_out);
        case 3: return launch512_v<OUT, WIN, true, false>(p, grid, stream, occ_out);
        default: return launch512_v<OUT, WIN, false, true>(p, grid, stream, occ_out);
    }
}

cudaError_t k1w512_launch(const K1WParams &p, int output, bool win, int mode, int grid, cudaStream_t stream, int *occ_out) {
    switch (output * 2 + (win ? 1 : 0)) {
        case 0: return launch512<OUT_COMPLEX, false>(p, mode, grid, stream, occ_out);
        case 1: return launch512<OUT_COMPLEX, true>(p, mode, grid, stream, occ_out);
        case 2: return launch512<OUT_MAG, false>(p, mode, grid, stream, occ_out);
        case 3: return launch512<OUT_MAG, true>(p, mode, grid, stream, occ_out);
        case 4: return launch512<OUT_POWER, false>(p, mode, grid, stream, occ_out);
        case 5: return launch512<OUT_POWER, true>(p, mode, grid, stream, occ_out);
        case 6: return launch512<OUT_DB, false>(p, mode, grid, stream, occ_out);
        case 7: return launch512<OUT_DB, true>(p, mode, grid, stream, occ_out);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace kopek
