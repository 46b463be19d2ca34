// K1M instantiations (fft_team.cuh): team-per-frame kernels for N = 2048 (R = 4) and 4096 (R = 8).  Compiled once per
// R (-DKOPEK_K1M_R=4 / 8, kopek_b200/build.py) so that the two sets of instantiations build in parallel; the host
// tables and the dispatcher live in the R = 4 object.
#include <atomic>
#include <cmath>

// Complex add / subtract as ONE packed instruction (add/sub.rn.f32x2 -> SASS FADD2) in this translation unit: the team
// kernels are bound by instruction issue, not by HBM, and 10 % fewer issue slots are worth +2.5 to +4 points of the HBM
// roofline at both sizes, burst and sustained (profiles/r2_sustained_sweep.txt).  Bit-identical results.
#define KOPEK_F32X2
#include "k1w_launch.cuh"
#include "fft_team.cuh"

#ifndef KOPEK_K1M_R
#error "compile with -DKOPEK_K1M_R=4 or 8"
#endif

namespace kopek {

#if KOPEK_K1M_R == 4
template <int R> static void build_tables_r(const float *half_window, double cs_a, double cs_b, float2 *out) {
    using G = K1MGeom<R>;
    constexpr int T = G::T, QD = G::QD;
    const double tau = 2.0 * M_PI;
    for (int i = 0; i < G::WIN_F2; ++i)
        out[i] = half_window ? make_float2(half_window[2 * i], half_window[2 * i + 1]) : make_float2(0.5f, 0.5f);
    float2 *rp = out + G::WIN_F2;
    for (int k1 = 0; k1 < 16; ++k1)
        for (int t = 0; t < T; ++t) {
            // W_M^{t k1} W_R^{(t % R)(t / R)}: both exponents reduced exactly before the division
            const int a = (t * k1) % G::M, b = ((t % R) * (t / R)) % R;
            const double x = -tau * ((double)a / G::M + (double)b / R);
            rp[k1 * T + t] = make_float2((float)cos(x), (float)sin(x));
        }
    float2 *tw = rp + G::RP_F2;
    for (int x = 0; x < R; ++x)
        for (int d = 0; d < QD; ++d)
            for (int t = 0; t < T; ++t) {
                const int r = t % R, rs = (r + x) % R, j = QD * r + d;
                const double a = -tau * (double)((rs * j) % T) / T;
                tw[(x * QD + d) * T + t] = make_float2((float)cos(a), (float)sin(a));
            }
    float2 *ph = tw + G::TW_F2;
    for (int q = 1; q <= R / 2; ++q)
        for (int t = 0; t < T; ++t) {
            const double a = -tau * (double)(((t % R) * q) % R) / R;
            ph[(q - 1) * T + t] = make_float2((float)cos(a), (float)sin(a));
        }
    float2 *pw = ph + G::PH_F2;
    for (int t = 0; t < T; ++t) {
        const double a = tau * (double)t / G::N;
        pw[t] = make_float2((float)-sin(a), (float)-cos(a));
    }
    // cosine-sum window w[n] = a - b cos(2 pi n / N), halved like the table: lane constants at n = 2 t + c
    float2 *cs = pw + G::PW_F2;
    for (int t = 0; t < T; ++t) {
        for (int c = 0; c < 2; ++c) {
            const double th = tau * (double)(2 * t + c) / G::N;
            cs[c * T + t] = make_float2((float)(-0.5 * cs_b * cos(th)), (float)(0.5 * cs_b * sin(th)));
        }
        cs[2 * T + t] = make_float2((float)(0.5 * cs_a), 0.0f);
        cs[3 * T + t] = out[t];                  // the frame edges keep their table values (n1 = 0, n1 = 15)
        cs[4 * T + t] = out[15 * T + t];
    }
}

int k1m_table_f4(int n) {
    switch (n) {
        case 2048: return K1MGeom<4>::TABLE_F2 / 2;
        case 4096: return K1MGeom<8>::TABLE_F2 / 2;
        default: return 0;
    }
}

void k1m_build_tables(int n, const float *half_window, double cs_a, double cs_b, float4 *out) {
    float2 *o = reinterpret_cast<float2 *>(out);
    if (n == 2048) build_tables_r<4>(half_window, cs_a, cs_b, o);
    else build_tables_r<8>(half_window, cs_a, cs_b, o);
}
#endif

#ifndef KOPEK_K1M_PF
#define KOPEK_K1M_PF 0      // L2 prefetch distance in frames per team: measured 0/2/3/5 -> no gain (not latency bound on HBM)
#endif
// resident teams per SM the register allocation is held to: 12 warps per SM (168 registers) -- measured 76 -> 86 %
// (N = 2048) and 70 -> 80 % (N = 4096) against 8 warps.  16 warps (-DKOPEK_K1M_WARPS=16: 128 registers,
// ~110 bytes of spills) is far worse: 84 -> 64 % (N = 2048) and 76 -> 57 % (N = 4096).
#ifndef KOPEK_K1M_WARPS
#define KOPEK_K1M_WARPS 12
#endif
template <int R, int WINK> struct K1MOcc {
    static constexpr int MINB = KOPEK_K1M_WARPS * 2 / R;
    static constexpr size_t SMEM = K1MGeom<R>::SMEM + (WINK == 3 ? K1MGeom<R>::SMEM_WIN : 0);
};

template <int R, int OUT, int WIN, bool FULL, bool VEC>
static cudaError_t launch_m(const K1WParams &p, int grid, cudaStream_t stream, int *occ_out) {
    auto kern = k1m_kernel<R, OUT, WIN, K1MOcc<R, WIN>::MINB, KOPEK_K1M_PF, FULL, VEC>;
    constexpr size_t smem = K1MOcc<R, WIN>::SMEM;
    if constexpr (smem > 48 * 1024) {
        // opt in to more than 48 KB of dynamic shared memory: once per instantiation and device (idempotent: a race is benign)
        static std::atomic<bool> opted[64];
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !opted[dev].load(std::memory_order_acquire)) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            if (dev >= 0 && dev < 64) opted[dev].store(true, std::memory_order_release);
        }
    }
    if (occ_out) {
        int nb = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 16 * R, smem);
        *occ_out = nb;                                       // teams (CTAs) per SM
        return e;
    }
    kern<<<grid, 16 * R, smem, stream>>>(p);
    return cudaGetLastError();
}

// mode: 0 = HALF rows, 64-bit loads; 1 = FULL rows; 2 = HALF rows, 32-bit loads; 3 = FULL rows, 32-bit loads
template <int R, int OUT, int WIN>
static cudaError_t launch_w(const K1WParams &p, int mode, int grid, cudaStream_t stream, int *occ_out) {
    if (mode == 1) return launch_m<R, OUT, WIN, true, true>(p, grid, stream, occ_out);
    if (mode == 2) return launch_m<R, OUT, WIN, false, false>(p, grid, stream, occ_out);
    if (mode == 3) return launch_m<R, OUT, WIN, true, false>(p, grid, stream, occ_out);
    return launch_m<R, OUT, WIN, false, true>(p, grid, stream, occ_out);
}

template <int R>
static cudaError_t launch_r(const K1WParams &p, int output, int wink, int mode, int grid, cudaStream_t stream,
                            int *occ_out) {
    switch (output * 4 + wink) {
        case 0: return launch_w<R, OUT_COMPLEX, 0>(p, mode, grid, stream, occ_out);
        case 2: return launch_w<R, OUT_COMPLEX, 2>(p, mode, grid, stream, occ_out);
        case 3: return launch_w<R, OUT_COMPLEX, 3>(p, mode, grid, stream, occ_out);
        case 4: return launch_w<R, OUT_MAG, 0>(p, mode, grid, stream, occ_out);
        case 6: return launch_w<R, OUT_MAG, 2>(p, mode, grid, stream, occ_out);
        case 7: return launch_w<R, OUT_MAG, 3>(p, mode, grid, stream, occ_out);
        case 8: return launch_w<R, OUT_POWER, 0>(p, mode, grid, stream, occ_out);
        case 10: return launch_w<R, OUT_POWER, 2>(p, mode, grid, stream, occ_out);
        case 11: return launch_w<R, OUT_POWER, 3>(p, mode, grid, stream, occ_out);
        case 12: return launch_w<R, OUT_DB, 0>(p, mode, grid, stream, occ_out);
        case 14: return launch_w<R, OUT_DB, 2>(p, mode, grid, stream, occ_out);
        case 15: return launch_w<R, OUT_DB, 3>(p, mode, grid, stream, occ_out);
        default: return cudaErrorInvalidValue;
    }
}

#if KOPEK_K1M_R == 4
cudaError_t k1m_launch_r8(const K1WParams &p, int output, int wink, int mode, int grid, cudaStream_t stream, int *occ_out);
cudaError_t k1m_launch(int n, const K1WParams &p, int output, int wink, int mode, int grid, cudaStream_t stream,
                       int *occ_out) {
    if (wink < 0 || wink > 3 || wink == 1 || mode < 0 || mode > 3) return cudaErrorInvalidValue;
    switch (n) {
        case 2048: return launch_r<4>(p, output, wink, mode, grid, stream, occ_out);
        case 4096: return k1m_launch_r8(p, output, wink, mode, grid, stream, occ_out);
        default: return cudaErrorInvalidValue;
    }
}
#else
cudaError_t k1m_launch_r8(const K1WParams &p, int output, int wink, int mode, int grid, cudaStream_t stream, int *occ_out) {
    return launch_r<8>(p, output, wink, mode, grid, stream, occ_out);
}
#endif

}  // namespace kopek
