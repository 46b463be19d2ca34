// K1W-512 -- 512-point real-input FFT + fused window / magnitude epilogue, SIXTEEN LANES per frame (sm_100a).
//
// Same design rules as fft_warp256.cuh / fft_warp1024.cuh: a lane group owns its frame end to end, every
// loop-invariant table (window slice, twiddles) lives in registers, the only shared-memory traffic is the
// conflict-free exchange, the next frames are prefetched into registers.  Two frames share a warp (lane = 16 g + a):
//
//   M = 256 = 16 x 16,  n = 16 n1 + n0,  k = k1 + 16 k0
//   A: lane a = n0 loads z[16 n1 + a] (n1 = 0..15), radix-16 over n1, multiplies by W256^{a k1}, publishes (k1, n0)
//   B: lane a = k1 gathers its 16 n0, radix-16 -> Z[a + 16 k0] in registers
//   split: the partner Z[M-k] lives in lane 16-a; only k0 >= 8 is published, each lane finishes 8 (k, M-k) pairs
// Element (k1, n0) sits at 64-bit position 17 k1 + n0 (one spare per 16), which makes both the publishing stores
// (consecutive lanes) and the gathering loads (stride 17) conflict free.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fft_warp256.cuh"

namespace kopek {

constexpr int K512_WIN_F2 = 256;           // 0.5 * w[2i], 0.5 * w[2i+1]
constexpr int K512_TW_F2 = 15 * 16;        // [k1-1][a] W256^{a k1}
constexpr int K512_PW_F2 = 16;             // [a] -i W512^{a}
constexpr int K512_TABLE_F2 = K512_WIN_F2 + K512_TW_F2 + K512_PW_F2;
constexpr int K512_FRAME_F2 = 16 * 17;     // exchange positions per frame
constexpr int K512_EX_F2 = 2 * K512_FRAME_F2;

// FULL / VEC: see fft_warp1024b.cuh (all N bins per row with the conjugate mirror stored beside the computed half;
// VEC == false reads frames that start at any 4-byte aligned sample with two 32-bit loads per pair).
template <int OUT, bool WIN, int WARPS, int MINB, bool FULL = false, bool VEC = true>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k1w_512_kernel(const K1WParams p) {
    constexpr int M = 256;
    extern __shared__ __align__(16) unsigned char k512_smem[];
    const int l = threadIdx.x & 31, g = l >> 4, a = l & 15;
    float2 *ex = reinterpret_cast<float2 *>(k512_smem) + (threadIdx.x >> 5) * K512_EX_F2 + K512_FRAME_F2 * g;
    float2 *exA = ex + a;                                   // + 17 k1
    const float2 *exB = ex + 17 * a;                        // + n0
    float2 *zW = ex + a;                                    // + 16 (k0 - 8)
    // partner lane 16 - a holds Z[M - k] at k0' = 15 - k0; lane 0 pairs with itself at k0' = 16 - k0 (one row further)
    const float2 *zR = ex + ((16 - a) & 15) + (a == 0 ? 16 : 0);   // + 16 (7 - k0)

    const float2 *tab = reinterpret_cast<const float2 *>(p.tables);
    float2 rw[16], rp[16];
    if constexpr (WIN) {
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) rw[n1] = tab[16 * n1 + a];
    }
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) rp[k1] = tab[K512_WIN_F2 + 16 * (k1 - 1) + a];
    const float2 pw_lane = tab[K512_WIN_F2 + K512_TW_F2 + a];      // -i W512^{a}

    const uint64_t n_groups = (p.n_frames + 1) >> 1;
    const uint64_t total_warps = (uint64_t)gridDim.x * WARPS;
    const uint64_t first = (uint64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
    auto load_group = [&](uint64_t grp, float2 *dst) {
        uint64_t fr = 2 * grp + g;
        if (fr >= p.n_frames) fr = p.n_frames - 1;
        if constexpr (VEC) {
            const float2 *xf = reinterpret_cast<const float2 *>(p.x + fr * p.hop) + a;
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) dst[n1] = __ldg(xf + 16 * n1);
        } else {
            const float *xs = p.x + fr * p.hop + 2 * a;
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) dst[n1] = make_float2(__ldg(xs + 32 * n1), __ldg(xs + 32 * n1 + 1));
        }
    };
    float2 vn[16];
    if (first < n_groups) load_group(first, vn);

    for (uint64_t grp = first; grp < n_groups; grp += total_warps) {
        const uint64_t frame = 2 * grp + g;
        const bool valid = frame < p.n_frames;
        float2 e[16];
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
            if constexpr (WIN) e[n1] = make_float2(vn[n1].x * rw[n1].x, vn[n1].y * rw[n1].y);
            else e[n1] = vn[n1];
        }
        if (grp + total_warps < n_groups) load_group(grp + total_warps, vn);

        // ---- A: radix-16 over n1, twiddle, publish (k1, n0 = a)
        Dft<16>::run(e);
        exA[0] = e[0];
#pragma unroll
        for (int k1 = 1; k1 < 16; ++k1) exA[17 * k1] = cmul(e[k1], rp[k1]);
        __syncwarp();
        // ---- B: lane a = k1 gathers n0 = 0..15, radix-16 -> Z[a + 16 k0]
#pragma unroll
        for (int n0 = 0; n0 < 16; ++n0) e[n0] = exB[n0];
        Dft<16>::run(e);
        __syncwarp();
        // ---- split: publish k0 >= 8, finish the pairs (k, M-k) for k0 < 8
#pragma unroll
        for (int k0 = 8; k0 < 16; ++k0) zW[16 * (k0 - 8)] = e[k0];
        __syncwarp();
        float2 b[8];
#pragma unroll
        for (int k0 = 0; k0 < 8; ++k0) b[k0] = zR[16 * (7 - k0)];
        if (a == 0) b[0] = e[0];                              // k = 0: Z[M] == Z[0]
        constexpr int EB = OUT == OUT_COMPLEX ? 8 : 4;
        char *o_lo = reinterpret_cast<char *>(p.out) + (frame * p.pitch + a) * EB;
        char *o_hi = reinterpret_cast<char *>(p.out) + (frame * p.pitch + (M - a)) * EB;
        char *m_lo = reinterpret_cast<char *>(p.out) + (frame * p.pitch + (2 * M - a)) * EB;   // mirrors (FULL): bin N - k
        char *m_hi = reinterpret_cast<char *>(p.out) + (frame * p.pitch + (M + a)) * EB;
        if (valid) {
#pragma unroll
            for (int k0 = 0; k0 < 8; ++k0) {
                const float2 A = e[k0], B = b[k0];
                float2 S = make_float2(A.x + B.x, A.y - B.y);
                float2 D = make_float2(A.x - B.x, A.y + B.y);
                float2 Q = cmul(mul_w32(D, k0), pw_lane);     // -i W512^{a + 16 k0} = (-i W512^a) W32^{k0}
                if constexpr (FULL) {
                    const bool edge = k0 == 0 && a == 0;      // bins 0 and M: no mirror
                    emit_pair<OUT, WIN>(o_lo + 16 * k0 * EB, edge ? nullptr : m_lo - 16 * k0 * EB, S + Q);
                    emit_pair<OUT, WIN>(o_hi - 16 * k0 * EB, edge ? nullptr : m_hi + 16 * k0 * EB,
                                        make_float2(S.x - Q.x, Q.y - S.y));
                } else {
                    emit_at<OUT, WIN>(o_lo + 16 * k0 * EB, S + Q);
                    emit_at<OUT, WIN>(o_hi - 16 * k0 * EB, make_float2(S.x - Q.x, Q.y - S.y));
                }
            }
            if (a == 0) {
                const float2 Xq = make_float2(2.0f * e[8].x, -2.0f * e[8].y);
                if constexpr (FULL) emit_pair<OUT, WIN>(o_lo + (M / 2) * EB, m_lo - (M / 2) * EB, Xq);
                else emit_at<OUT, WIN>(o_lo + (M / 2) * EB, Xq);
            }
        }
        __syncwarp();
    }
}

}  // namespace kopek
