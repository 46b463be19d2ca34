// K1W-256 -- 256-point real-input FFT + fused window / magnitude epilogue, EIGHT LANES per frame (sm_100a).
//
// BASELINE.json configs[2]: "1M independent 256-point FFTs of white noise ... frame-sharded across 1/2/4/8 B200".
// Same design rules as the 1024-point kernel (fft_warp1024.cuh): a lane group owns its frame end to end, every
// loop-invariant table lives in registers, the only shared-memory traffic is the 128-bit conflict-free exchange,
// the next frames are prefetched into registers.  Four frames share a warp (lane = 8 g + a):
//
//   M = 128 = 8 x 16,  n = 16 n1 + n0,  k = k1 + 8 k0
//   A: lane a holds the pair n0 = 2a, 2a+1; radix-8 over n1; multiply by W128^{n0 k1}; publish (k1, n0 pair)
//   B: lane a = k1 reads all 16 n0 (8 float4), radix-16 -> Z[a + 8 k0] in registers
//   split: the partner Z[M-k] lives in lane 8-a; only k0 >= 8 is published, each lane finishes 8 (k, M-k) pairs
// One spare float4 per eight (slot (k1, s) at 9 k1 + s) makes both exchange patterns conflict free;
// tools/k1w256_model.py is the numpy model of these address functions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "lane_common.cuh"

namespace kopek {

constexpr int K256_WIN_F4 = 64;            // 0.5 * w[4i .. 4i+3]
constexpr int K256_TW_F4 = 7 * 8;          // [k1-1][a] (W128^{2a k1}, W128^{(2a+1) k1})
constexpr int K256_PW_F4 = 8;              // [a] (-i W256^{a}, unused)
constexpr int K256_TABLE_F4 = K256_WIN_F4 + K256_TW_F4 + K256_PW_F4;
constexpr int K256_EX_F4 = 288;            // per warp: 4 frames x 72 slots

// multiply by W32^k0 = W256^{8 k0} (compile-time part of the post twiddle -i W256^{a + 8 k0})
__device__ __forceinline__ float2 mul_w32(float2 v, int k0) {
    constexpr float c[8] = {1.0f, 0.98078528040323043058f, 0.92387953251128673848f, 0.83146961230254523567f,
                            0.70710678118654757274f, 0.55557023301960228867f, 0.38268343236508983729f,
                            0.19509032201612833135f};
    constexpr float s[8] = {0.0f, -0.19509032201612824808f, -0.38268343236508978178f, -0.55557023301960217765f,
                            -0.70710678118654746172f, -0.83146961230254523567f, -0.92387953251128673848f,
                            -0.98078528040323043058f};
    return k0 == 0 ? v : cmul_const(v, c[k0], s[k0]);
}

// FULL / VEC: see fft_warp1024b.cuh (all N bins per row with the conjugate mirror stored beside the computed half;
// VEC == false reads frames that start at any 4-byte aligned sample with four 32-bit loads per quad).
template <int OUT, bool WIN, int WARPS, int MINB, bool FULL = false, bool VEC = true>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k1w_256_kernel(const K1WParams p) {
    constexpr int M = 128;
    extern __shared__ __align__(16) unsigned char k256_smem[];
    const int l = threadIdx.x & 31, g = l >> 3, a = l & 7;
    float4 *ex = reinterpret_cast<float4 *>(k256_smem) + (threadIdx.x >> 5) * K256_EX_F4 + 72 * g;
    float4 *exA = ex + a;                          // + 9 k1
    const float4 *exB = ex + 9 * a;                // + s
    float4 *zW = ex + a;                           // + 8 j      (k0 = 8 + 2j, 9 + 2j)
    const float4 *zR = ex + ((8 - a) & 7);         // + 8 (3 - i)

    // loop-invariant tables -> registers
    float4 rw[8];
    float2 rp0[8], rp1[8];
    if constexpr (WIN) {
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) rw[n1] = p.tables[8 * n1 + a];
    }
#pragma unroll
    for (int k1 = 1; k1 < 8; ++k1) {
        float4 t = p.tables[K256_WIN_F4 + 8 * (k1 - 1) + a];
        rp0[k1] = make_float2(t.x, t.y);
        rp1[k1] = make_float2(t.z, t.w);
    }
    const float4 pwt = p.tables[K256_WIN_F4 + K256_TW_F4 + a];
    const float2 pw_lane = make_float2(pwt.x, pwt.y);      // -i W256^{a}

    const uint64_t n_groups = (p.n_frames + 3) >> 2;
    const uint64_t total_warps = (uint64_t)gridDim.x * WARPS;
    const uint64_t first = (uint64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
    auto load_group = [&](uint64_t grp, float4 *dst) {
        uint64_t fr = 4 * grp + g;
        if (fr >= p.n_frames) fr = p.n_frames - 1;
        if constexpr (VEC) {
            const float4 *xf = reinterpret_cast<const float4 *>(p.x + fr * p.hop) + a;
#pragma unroll
            for (int n1 = 0; n1 < 8; ++n1) dst[n1] = ldg_stream(xf + 8 * n1);
        } else {
            const float *xs = p.x + fr * p.hop + 4 * a;
#pragma unroll
            for (int n1 = 0; n1 < 8; ++n1)
                dst[n1] = make_float4(__ldg(xs + 32 * n1), __ldg(xs + 32 * n1 + 1), __ldg(xs + 32 * n1 + 2),
                                      __ldg(xs + 32 * n1 + 3));
        }
    };
    float4 vn[8];
    if (first < n_groups) load_group(first, vn);

    for (uint64_t grp = first; grp < n_groups; grp += total_warps) {
        const uint64_t frame = 4 * grp + g;
        const bool valid = frame < p.n_frames;
        float4 v[8];
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) v[n1] = vn[n1];
        if (grp + total_warps < n_groups) load_group(grp + total_warps, vn);

        // ---- A: window, radix-8 over n1, twiddle, publish
        float2 d0[8], d1[8];
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
            if constexpr (WIN) {
                d0[n1] = make_float2(v[n1].x * rw[n1].x, v[n1].y * rw[n1].y);
                d1[n1] = make_float2(v[n1].z * rw[n1].z, v[n1].w * rw[n1].w);
            } else {
                d0[n1] = make_float2(v[n1].x, v[n1].y);
                d1[n1] = make_float2(v[n1].z, v[n1].w);
            }
        }
        Dft<8>::run(d0);
        Dft<8>::run(d1);
        exA[0] = make_float4(d0[0].x, d0[0].y, d1[0].x, d1[0].y);
#pragma unroll
        for (int k1 = 1; k1 < 8; ++k1) {
            float2 x0 = cmul(d0[k1], rp0[k1]), x1 = cmul(d1[k1], rp1[k1]);
            exA[9 * k1] = make_float4(x0.x, x0.y, x1.x, x1.y);
        }
        __syncwarp();
        // ---- B: lane a = k1 gathers its 16 n0, radix-16 -> Z[a + 8 k0]
        float2 e[16];
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            float4 t = exB[s];
            e[2 * s] = make_float2(t.x, t.y);
            e[2 * s + 1] = make_float2(t.z, t.w);
        }
        Dft<16>::run(e);
        __syncwarp();
        // ---- split: publish k0 >= 8, finish the pairs (k, M-k) for k0 < 8
#pragma unroll
        for (int j = 0; j < 4; ++j) zW[8 * j] = make_float4(e[8 + 2 * j].x, e[8 + 2 * j].y, e[9 + 2 * j].x, e[9 + 2 * j].y);
        __syncwarp();
        float2 b[8];                       // partner Z[M-k] of bin k = a + 8 k0
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 t = zR[8 * (3 - i)];
            b[2 * i] = make_float2(t.z, t.w);
            b[2 * i + 1] = a ? make_float2(t.x, t.y) : make_float2(t.z, t.w);
        }
        if (a == 0) {                      // k1 = 0 lanes: even k0 pair with slot (0, 4 - i).lo; k = 0 with itself
            b[0] = e[0];
#pragma unroll
            for (int i = 1; i < 4; ++i) {
                float4 t = ex[8 * (4 - i)];
                b[2 * i] = make_float2(t.x, t.y);
            }
        }
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
                float2 Q = cmul(mul_w32(D, k0), pw_lane);
                if constexpr (FULL) {
                    const bool edge = k0 == 0 && a == 0;      // bins 0 and M: no mirror
                    emit_pair<OUT, WIN>(o_lo + 8 * k0 * EB, edge ? nullptr : m_lo - 8 * k0 * EB, S + Q);
                    emit_pair<OUT, WIN>(o_hi - 8 * k0 * EB, edge ? nullptr : m_hi + 8 * k0 * EB,
                                        make_float2(S.x - Q.x, Q.y - S.y));
                } else {
                    emit_at<OUT, WIN>(o_lo + 8 * k0 * EB, S + Q);
                    emit_at<OUT, WIN>(o_hi - 8 * k0 * EB, make_float2(S.x - Q.x, Q.y - S.y));
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
