// Instantiations + host launchers of the 1024-point lane-group kernel (fft_warp1024b.cuh); the 256- and 512-point
// kernels are instantiated in k1w_small_inst.cu so that the two sets compile in parallel.
#include <math.h>

#include <vector>

#include "k1w_launch.cuh"
#include "fft_warp1024b.cuh"

namespace kopek {

// ------------------------------------------------------------------ 1024-point, two-exchange kernel ----
void k1b_build_tables(const float *half_window /* 1024 or nullptr */, float4 *out4) {
    float2 *out = reinterpret_cast<float2 *>(out4);
    const double tau = 2.0 * M_PI;
    for (int i = 0; i < K1B_WIN_F2; ++i)
        out[i] = half_window ? make_float2(half_window[2 * i], half_window[2 * i + 1]) : make_float2(0.5f, 0.5f);
    float2 *tw = out + K1B_WIN_F2;
    for (int k = 1; k < 16; ++k)
        for (int l = 0; l < 32; ++l) {
            double x = -tau * (double)(l * k) / 512.0, sg = (l & 3) == 3 ? -1.0 : 1.0;
            tw[(k - 1) * 32 + l] = make_float2((float)(sg * cos(x)), (float)(sg * sin(x)));
        }
    float2 *cw = tw + K1B_TW_F2;
    for (int k = 0; k < 8; ++k)
        for (int par = 0; par < 2; ++par) {
            double x = -tau * (double)(k + 8 * par) / 32.0;
            cw[2 * k + par] = make_float2((float)cos(x), (float)sin(x));
        }
    float2 *pw = cw + K1B_CW_F2;
    for (int l = 0; l < 32; ++l) {
        double x = tau * (double)((l >> 1) + 128 * (l & 1)) / 1024.0;
        pw[l] = make_float2((float)-sin(x), (float)-cos(x));
    }
}

template <int OUT, bool WIN, bool BULK, bool PCM, bool FULL, bool VEC>
static cudaError_t launch1b_v(const K1WParams &p, int grid, cudaStream_t stream, int *occ_out) {
    constexpr int WARPS = 4;
    auto kern = k1w_1024b_kernel<OUT, WIN, WARPS, 3, BULK, PCM, FULL, VEC>;
    constexpr size_t smem = sizeof(float2) * WARPS * (K1B_EX_F2 + (BULK ? K1B_STAGE_F2 : 0));
    if (occ_out) {
        int nb = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, WARPS * 32, smem);
        *occ_out = nb * WARPS;
        return e;
    }
    kern<<<grid, WARPS * 32, smem, stream>>>(p);
    return cudaGetLastError();
}

// p.vec picks the instantiation: 64-bit loads, or two 32-bit loads per pair for odd hops / unaligned bases (spectra
// written directly only: the bulk-store and PCM forms keep their alignment requirements)
template <int OUT, bool WIN, bool BULK = false, bool PCM = false, bool FULL = false>
static cudaError_t launch1b(const K1WParams &p, int grid, cudaStream_t stream, int *occ_out) {
    if constexpr (!BULK && !PCM) {
        if (!p.vec && !occ_out) return launch1b_v<OUT, WIN, false, false, FULL, false>(p, grid, stream, occ_out);
    }
    return launch1b_v<OUT, WIN, BULK, PCM, FULL, true>(p, grid, stream, occ_out);
}

cudaError_t k1b_launch_pcm(const K1WParams &p, int output, bool win, int grid, cudaStream_t stream) {
    switch (output * 2 + (win ? 1 : 0)) {
        case 0: return launch1b<OUT_COMPLEX, false, false, true>(p, grid, stream, nullptr);
        case 1: return launch1b<OUT_COMPLEX, true, false, true>(p, grid, stream, nullptr);
        case 2: return launch1b<OUT_MAG, false, false, true>(p, grid, stream, nullptr);
        case 3: return launch1b<OUT_MAG, true, false, true>(p, grid, stream, nullptr);
        case 4: return launch1b<OUT_POWER, false, false, true>(p, grid, stream, nullptr);
        case 5: return launch1b<OUT_POWER, true, false, true>(p, grid, stream, nullptr);
        case 6: return launch1b<OUT_DB, false, false, true>(p, grid, stream, nullptr);
        case 7: return launch1b<OUT_DB, true, false, true>(p, grid, stream, nullptr);
        case 8: return launch1b<OUT_BANDS_LOG, false, false, true>(p, grid, stream, nullptr);
        case 9: return launch1b<OUT_BANDS_LOG, true, false, true>(p, grid, stream, nullptr);
        case 10: return launch1b<OUT_BANDS_LOW, false, false, true>(p, grid, stream, nullptr);
        case 11: return launch1b<OUT_BANDS_LOW, true, false, true>(p, grid, stream, nullptr);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t k1b_launch_full(const K1WParams &p, int output, bool win, int grid, cudaStream_t stream) {
    switch (output * 2 + (win ? 1 : 0)) {
        case 0: return launch1b<OUT_COMPLEX, false, false, false, true>(p, grid, stream, nullptr);
        case 1: return launch1b<OUT_COMPLEX, true, false, false, true>(p, grid, stream, nullptr);
        case 2: return launch1b<OUT_MAG, false, false, false, true>(p, grid, stream, nullptr);
        case 3: return launch1b<OUT_MAG, true, false, false, true>(p, grid, stream, nullptr);
        case 4: return launch1b<OUT_POWER, false, false, false, true>(p, grid, stream, nullptr);
        case 5: return launch1b<OUT_POWER, true, false, false, true>(p, grid, stream, nullptr);
        case 6: return launch1b<OUT_DB, false, false, false, true>(p, grid, stream, nullptr);
        case 7: return launch1b<OUT_DB, true, false, false, true>(p, grid, stream, nullptr);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t k1b_launch(const K1WParams &p, int output, bool win, bool bulk, int grid, cudaStream_t stream,
                       int *occ_out) {
    if (bulk) {
        switch (output * 2 + (win ? 1 : 0)) {
            case 2: return launch1b<OUT_MAG, false, true>(p, grid, stream, occ_out);
            case 3: return launch1b<OUT_MAG, true, true>(p, grid, stream, occ_out);
            case 4: return launch1b<OUT_POWER, false, true>(p, grid, stream, occ_out);
            case 5: return launch1b<OUT_POWER, true, true>(p, grid, stream, occ_out);
            case 6: return launch1b<OUT_DB, false, true>(p, grid, stream, occ_out);
            case 7: return launch1b<OUT_DB, true, true>(p, grid, stream, occ_out);
            default: return cudaErrorInvalidValue;
        }
    }
    switch (output * 2 + (win ? 1 : 0)) {
        case 0: return launch1b<OUT_COMPLEX, false>(p, grid, stream, occ_out);
        case 1: return launch1b<OUT_COMPLEX, true>(p, grid, stream, occ_out);
        case 2: return launch1b<OUT_MAG, false>(p, grid, stream, occ_out);
        case 3: return launch1b<OUT_MAG, true>(p, grid, stream, occ_out);
        case 4: return launch1b<OUT_POWER, false>(p, grid, stream, occ_out);
        case 5: return launch1b<OUT_POWER, true>(p, grid, stream, occ_out);
        case 6: return launch1b<OUT_DB, false>(p, grid, stream, occ_out);
        case 7: return launch1b<OUT_DB, true>(p, grid, stream, occ_out);
        case 8: return launch1b<OUT_BANDS_LOG, false>(p, grid, stream, occ_out);
        case 9: return launch1b<OUT_BANDS_LOG, true>(p, grid, stream, occ_out);
        case 10: return launch1b<OUT_BANDS_LOW, false>(p, grid, stream, occ_out);
        case 11: return launch1b<OUT_BANDS_LOW, true>(p, grid, stream, occ_out);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace kopek
