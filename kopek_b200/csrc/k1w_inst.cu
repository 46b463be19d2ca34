// Instantiations + host launchers of the lane-group kernels: 1024 points (fft_warp1024b.cuh; the three-pass
// fft_warp1024.cuh in tuning builds), 256 points (fft_warp256.cuh), 512 points (fft_warp512.cuh).
#include <math.h>

#include <vector>

#include "k1w_launch.cuh"
#include "fft_warp1024b.cuh"

namespace kopek {

#ifdef KOPEK_K1W_SWEEP   // the three-pass kernel of fft_warp1024.cuh: tuning builds only (tools/sweep_k1w.py)
// Host-side tables, built in f64 and rounded once to f32 (layout: see fft_warp1024.cuh).
void k1w_build_tables(const float *half_window /* 1024 or nullptr */, float4 *out) {
    const double tau = 2.0 * M_PI;
    for (int i = 0; i < K1W_WIN_F4; ++i)
        out[i] = half_window ? make_float4(half_window[4 * i], half_window[4 * i + 1], half_window[4 * i + 2],
                                           half_window[4 * i + 3])
                             : make_float4(0.5f, 0.5f, 0.5f, 0.5f);
    float4 *tpp = out + K1W_WIN_F4;
    for (int k2 = 1; k2 < 8; ++k2)
        for (int l = 0; l < 32; ++l) {
            double a = -tau * (double)(l * k2) / 256.0, b = -tau * (double)((2 * l + 1) * k2) / 512.0;
            tpp[(k2 - 1) * 32 + l] = make_float4((float)cos(a), (float)sin(a), (float)cos(b), (float)sin(b));
        }
    float4 *tq = tpp + K1W_TPP_F4;
    for (int k1 = 1; k1 < 8; ++k1)
        for (int b = 0; b < 4; ++b) {
            double a = -tau * (double)(b * k1) / 32.0, c = -tau * (double)((2 * b + 1) * k1) / 64.0;
            tq[(k1 - 1) * 4 + b] = make_float4((float)cos(a), (float)sin(a), (float)cos(c), (float)sin(c));
        }
    float4 *pw = tq + K1W_TQ_F4;
    for (int k0 = 0; k0 < 4; ++k0)
        for (int l = 0; l < 32; ++l) {
            int k = (l >> 2) + 16 * (l & 3) + 64 * k0;
            double a = tau * (double)k / 1024.0, b = tau * (double)(k + 8) / 1024.0;
            pw[k0 * 32 + l] = make_float4((float)-sin(a), (float)-cos(a), (float)-sin(b), (float)-cos(b));
        }
}

template <int OUT, bool WIN, int WARPS, int MINB, int TWREG, bool WINREG, int LOAD, bool BULK = false>
static cudaError_t launch_v(const K1WParams &p, int grid, cudaStream_t stream, int *occ_out) {
    using G = K1WGeom<WARPS, LOAD == 1, BULK>;
    auto kern = k1w_1024_kernel<OUT, WIN, WARPS, MINB, TWREG, WINREG, LOAD, BULK>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
    if (e != cudaSuccess) return e;
    if (occ_out) {
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, G::BLOCK, G::SMEM);
        *occ_out = nb * WARPS;     // resident warps (= frames in flight) per SM
        return e;
    }
    kern<<<grid, G::BLOCK, G::SMEM, stream>>>(p);
    return cudaGetLastError();
}

// variant table: (warps per CTA, min CTAs per SM, twiddle registers 0/1/2, window in registers, load mode:
// 0 LDG, 1 TMA prefetch, 2 LDG register prefetch).  Variant 0 is the production configuration (measured best
// on B200, profiles/r1_k1w_sweep.md); the others are only compiled with -DKOPEK_K1W_SWEEP for tools/sweep_k1w.py.
//   windowed:    28 twiddle registers + 32 window registers + 32 prefetch registers (LDG of the next frame)
//   rectangular: 56 twiddle registers + 32 prefetch registers
// The TMA landing buffer (load mode 1) costs 27 shared-memory wavefronts per frame on the pipe that bounds the
// kernel, so it lost to the register prefetch once the register budget allowed both (90.9 % vs 86.9 % burst,
// 76.9 % vs 73.9 % sustained, Hann).
#ifdef KOPEK_K1W_SWEEP
#define KOPEK_K1W_VARIANTS(X)                                                                      \
    X(0, 4, 3, 1, true, 2) X(1, 4, 3, 2, true, 1) X(2, 4, 3, 1, true, 1) X(3, 4, 2, 1, true, 2)    \
    X(4, 4, 3, 2, true, 2) X(5, 4, 4, 1, true, 1) X(6, 4, 4, 0, true, 2) X(7, 4, 3, 0, true, 2)
#else
#define KOPEK_K1W_VARIANTS(X) X(0, 4, 3, 1, true, 2)
#endif

template <int OUT, bool WIN>
static cudaError_t launch_o(const K1WParams &p, int variant, int grid, cudaStream_t stream, int *occ_out) {
    if constexpr (OUT == OUT_MAG || OUT == OUT_POWER || OUT == OUT_DB) {
        if (variant == 100) {      // production shape with the bulk-store epilogue (kopek_plan_set_bulk_store)
            if (WIN) return launch_v<OUT, WIN, 4, 3, 1, true, 2, true>(p, grid, stream, occ_out);
            return launch_v<OUT, WIN, 4, 3, 2, true, 2, true>(p, grid, stream, occ_out);
        }
    }
    if (variant == 0 && !WIN) return launch_v<OUT, WIN, 4, 3, 2, true, 2>(p, grid, stream, occ_out);
    switch (variant) {
#define X(ID, W, B, T, R, A) case ID: return launch_v<OUT, WIN, W, B, T, R, A>(p, grid, stream, occ_out);
        KOPEK_K1W_VARIANTS(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
}

int k1w_variant_warps(int variant) {
    if (variant == 100) return 4;
    switch (variant) {
#define X(ID, W, B, T, R, A) case ID: return W;
        KOPEK_K1W_VARIANTS(X)
#undef X
        default: return 0;
    }
}

cudaError_t k1w_launch(const K1WParams &p, int output, bool win, int variant, int grid, cudaStream_t stream,
                       int *occ_out) {
    switch (output * 2 + (win ? 1 : 0)) {
        case 0: return launch_o<OUT_COMPLEX, false>(p, variant, grid, stream, occ_out);
        case 1: return launch_o<OUT_COMPLEX, true>(p, variant, grid, stream, occ_out);
        case 2: return launch_o<OUT_MAG, false>(p, variant, grid, stream, occ_out);
        case 3: return launch_o<OUT_MAG, true>(p, variant, grid, stream, occ_out);
        case 4: return launch_o<OUT_POWER, false>(p, variant, grid, stream, occ_out);
        case 5: return launch_o<OUT_POWER, true>(p, variant, grid, stream, occ_out);
        case 6: return launch_o<OUT_DB, false>(p, variant, grid, stream, occ_out);
        case 7: return launch_o<OUT_DB, true>(p, variant, grid, stream, occ_out);
        case 8: return launch_o<OUT_BANDS_LOG, false>(p, variant, grid, stream, occ_out);
        case 9: return launch_o<OUT_BANDS_LOG, true>(p, variant, grid, stream, occ_out);
        case 10: return launch_o<OUT_BANDS_LOW, false>(p, variant, grid, stream, occ_out);
        default: return launch_o<OUT_BANDS_LOW, true>(p, variant, grid, stream, occ_out);
    }
}

#endif  // KOPEK_K1W_SWEEP

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

template <int OUT, bool WIN, bool BULK = false, bool PCM = false>
static cudaError_t launch1b(const K1WParams &p, int grid, cudaStream_t stream, int *occ_out) {
    constexpr int WARPS = 4;
    auto kern = k1w_1024b_kernel<OUT, WIN, WARPS, 3, BULK, PCM>;
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

template <int OUT, bool WIN>
static cudaError_t launch256(const K1WParams &p, int grid, cudaStream_t stream, int *occ_out) {
    constexpr int WARPS = 4;
    auto kern = k1w_256_kernel<OUT, WIN, WARPS, 3>;
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

cudaError_t k1w256_launch(const K1WParams &p, int output, bool win, int grid, cudaStream_t stream, int *occ_out) {
    switch (output * 2 + (win ? 1 : 0)) {
        case 0: return launch256<OUT_COMPLEX, false>(p, grid, stream, occ_out);
        case 1: return launch256<OUT_COMPLEX, true>(p, grid, stream, occ_out);
        case 2: return launch256<OUT_MAG, false>(p, grid, stream, occ_out);
        case 3: return launch256<OUT_MAG, true>(p, grid, stream, occ_out);
        case 4: return launch256<OUT_POWER, false>(p, grid, stream, occ_out);
        case 5: return launch256<OUT_POWER, true>(p, grid, stream, occ_out);
        case 6: return launch256<OUT_DB, false>(p, grid, stream, occ_out);
        case 7: return launch256<OUT_DB, true>(p, grid, stream, occ_out);
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

template <int OUT, bool WIN>
static cudaError_t launch512(const K1WParams &p, int grid, cudaStream_t stream, int *occ_out) {
    constexpr int WARPS = 4;
    auto kern = k1w_512_kernel<OUT, WIN, WARPS, 3>;
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

cudaError_t k1w512_launch(const K1WParams &p, int output, bool win, int grid, cudaStream_t stream, int *occ_out) {
    switch (output * 2 + (win ? 1 : 0)) {
        case 0: return launch512<OUT_COMPLEX, false>(p, grid, stream, occ_out);
        case 1: return launch512<OUT_COMPLEX, true>(p, grid, stream, occ_out);
        case 2: return launch512<OUT_MAG, false>(p, grid, stream, occ_out);
        case 3: return launch512<OUT_MAG, true>(p, grid, stream, occ_out);
        case 4: return launch512<OUT_POWER, false>(p, grid, stream, occ_out);
        case 5: return launch512<OUT_POWER, true>(p, grid, stream, occ_out);
        case 6: return launch512<OUT_DB, false>(p, grid, stream, occ_out);
        case 7: return launch512<OUT_DB, true>(p, grid, stream, occ_out);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace kopek
