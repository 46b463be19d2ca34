// K1W-1024b -- 1024-point real-input FFT + fused window / magnitude epilogue in TWO shared-memory exchanges (sm_100a).
//
// The three-pass kernel (fft_warp1024.cuh) needs 164 exchange wavefronts per frame on the LSU data pipe, the
// pipe that bounds it (ncu: 83 % busy).  The 256- and 512-point kernels reach 95-99 % of the HBM roofline with a
// two-pass 16-lane/8-lane layout; this kernel carries that structure to one warp per 1024-point frame:
//
//   M = 512 = 16 x 32,  n = 32 n1 + n0,  k = k1 + 16 k0
//   A: lane l = n0 loads z[32 n1 + l] (n1 = 0..15), radix-16 over n1, multiplies by W512^{l k1}, publishes (k1, n0)
//   B: the 32-point transform over n0 of butterfly k1 is shared by the lane PAIR (k1 = l/2, par = l%2): each lane
//      gathers the 16 n0 of its parity and runs a radix-16 (even lane: E[j], odd lane: O[j]); the final radix-2,
//      X[j] = E[j] + W32^j O[j], X[j+16] = E[j] - W32^j O[j], needs only HALF of the partner's values: one
//      shfl.xor per word exchanges E[8..15] against O[0..7].  Lanes with n0 = 3 (mod 4) publish their pass-A
//      outputs negated, which rotates the odd lane's radix-16 output by 8, so both lanes send register [8 + k]
//      and no select is needed on the sending side.
//   split: the partner Z[M-k] of bin k = k1 + 16 k0 lives in lane pair 16 - k1, other parity; each lane keeps
//      eight bins k0 < 16 and publishes eight bins k0 >= 16.
// Shared-memory wavefronts per frame: 32 + 32 (exchange) + 16 (shuffles) + 16 + 16 (split) = 112 instead of 164;
// element (k1, n0) sits at 64-bit position 34 k1 + n0, which makes every access conflict free.
// tools/k1w1024v2_model.py is the numpy model of these address functions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "lane_common.cuh"

namespace kopek {

constexpr int K1B_WIN_F2 = 512;            // (0.5 w[2i], 0.5 w[2i+1])
constexpr int K1B_TW_F2 = 15 * 32;         // [k1-1][l]  +-W512^{l k1}  (negated for l = 3 mod 4)
constexpr int K1B_CW_F2 = 8 * 2;           // [k][par]   W32^{k + 8 par}
constexpr int K1B_PW_F2 = 32;              // [l]        -i W1024^{(l/2) + 128 (l%2)}
constexpr int K1B_TABLE_F2 = K1B_WIN_F2 + K1B_TW_F2 + K1B_CW_F2 + K1B_PW_F2;
constexpr int K1B_EX_F2 = 16 * 34;         // exchange positions per warp (4352 bytes)
constexpr int K1B_STAGE_F2 = 260;          // bulk-store epilogue: the frame's 513 f32 (+ pad to 16 bytes) per warp

__device__ __forceinline__ uint32_t k1b_smem_u32(const void *ptr) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(ptr));
}

// BULK (f32 outputs, rows 16-byte aligned: pitch % 4 == 0): the 513 values of a frame are staged in shared memory
// and leave as ONE cp.async.bulk store of 2048 bytes (+ bin 512) -- the form that reaches link bandwidth when
// p.out is a peer GPU's memory (kopek_plan_set_bulk_store).
// PCM: p.x points at interleaved STEREO i16 frames (kopek::decoder::decode, src/decoder.rs:1-31) and hop counts frames;
// the marshalling of examples/graph/player.rs:56-59 -- frame[channel] as f64 / i16::MAX, narrowed once to f32 -- is
// fused into the load stage, so decoded audio reaches its spectrum in one pass (4 bytes in per sample, like f32 mono).
// s / 32767 is evaluated as q = s r, q += fma(-q, 32767, s) r with r = RN(1/32767): for all 65536 inputs this equals
// the f64 quotient narrowed to f32 (tests/test_oracle.py::test_pcm16_marshalling checks the identity exhaustively).
__device__ __forceinline__ float pcm16_sample(float raw_bits, uint32_t shift) {
    const float sf = (float)((__float_as_int(raw_bits) << shift) >> 16);
    constexpr float r = 1.0f / 32767.0f;
    const float q = sf * r;
    return fmaf(fmaf(-q, 32767.0f, sf), r, q);
}

// FULL: all n = 1024 bins per row, as the reference's callers iterate them (examples/graph/utils.rs:47-63): bins
// 513..1023 are the conjugate mirror X[n - k] = conj X[k] of a real frame, so every value is computed once and stored
// twice (the mirror of a 64-byte run of bins is a 64-byte run again).
// VEC == false: frames start at any 4-byte aligned sample (odd hop, unaligned base): the 64-bit pair loads become two
// 32-bit loads each -- the same bytes over the same wavefronts, sixteen more load instructions per frame.
template <int OUT, bool WIN, int WARPS, int MINB, bool BULK = false, bool PCM = false, bool FULL = false, bool VEC = true>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k1w_1024b_kernel(const K1WParams p) {
    constexpr int M = 512;
    static_assert(!BULK || (OUT == OUT_MAG || OUT == OUT_POWER || OUT == OUT_DB), "bulk store: f32 spectra only");
    static_assert(!FULL || (!BULK && OUT <= OUT_DB), "full rows: spectra written directly");
    constexpr int WARP_F2 = K1B_EX_F2 + (BULK ? K1B_STAGE_F2 : 0);
    extern __shared__ __align__(16) unsigned char k1b_smem[];
    const int l = threadIdx.x & 31, k1 = l >> 1, par = l & 1;
    float2 *ex = reinterpret_cast<float2 *>(k1b_smem) + (threadIdx.x >> 5) * WARP_F2;
    float *stage = reinterpret_cast<float *>(ex + K1B_EX_F2);
    float2 *exA = ex + l;                                   // + 34 k1'
    const float2 *exB = ex + 34 * k1 + par;                 // + 2 m
    float2 *zW = ex + l;                                    // + 32 i   (upper set, index i)
    // partner of bin k1 + 16 k0 (k0 = 8 par + i): lane pair 16 - k1, other parity, upper index 7 - i;
    // k1 = 0: the other lane of the same pair, upper index 8 - i (one row further; i = 0 handled from registers)
    const float2 *zR = ex + (k1 ? 2 * (16 - k1) + (1 - par) : (l ^ 1) + 32);   // + 32 (7 - i)

    const float2 *tab = reinterpret_cast<const float2 *>(p.tables);
    float2 rw[16], rp[16], rc[8];
    if constexpr (WIN) {
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) rw[n1] = tab[32 * n1 + l];
    }
#pragma unroll
    for (int k = 1; k < 16; ++k) rp[k] = tab[K1B_WIN_F2 + 32 * (k - 1) + l];
#pragma unroll
    for (int k = 0; k < 8; ++k) rc[k] = tab[K1B_WIN_F2 + K1B_TW_F2 + 2 * k + par];
    const float2 pw_lane = tab[K1B_WIN_F2 + K1B_TW_F2 + K1B_CW_F2 + l];
    const float sgn0 = (l & 3) == 3 ? -1.0f : 1.0f;         // sign of the untwiddled k1' = 0 output

    const uint64_t total_warps = (uint64_t)gridDim.x * WARPS;
    const uint64_t first = (uint64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
    auto load_frame = [&](uint64_t fr, float2 *dst) {
        if constexpr (PCM || VEC) {
            const float2 *xf = reinterpret_cast<const float2 *>(p.x + fr * p.hop) + l;
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) dst[n1] = __ldg(xf + 32 * n1);
        } else {
            const float *xs = p.x + fr * p.hop + 2 * l;
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) dst[n1] = make_float2(__ldg(xs + 64 * n1), __ldg(xs + 64 * n1 + 1));
        }
    };
    float2 vn[16];
    if (first < p.n_frames) load_frame(first, vn);

    for (uint64_t frame = first; frame < p.n_frames; frame += total_warps) {
        float2 e[16];
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
            float2 v = vn[n1];
            if constexpr (PCM) v = make_float2(pcm16_sample(v.x, p.pcm_shift), pcm16_sample(v.y, p.pcm_shift));
            if constexpr (WIN) e[n1] = make_float2(v.x * rw[n1].x, v.y * rw[n1].y);
            else e[n1] = v;
        }
        if (frame + total_warps < p.n_frames) load_frame(frame + total_warps, vn);

        // ---- A: radix-16 over n1, twiddle (sign folded in), publish (k1', n0 = l)
        Dft<16>::run(e);
        exA[0] = make_float2(e[0].x * sgn0, e[0].y * sgn0);
#pragma unroll
        for (int k = 1; k < 16; ++k) exA[34 * k] = cmul(e[k], rp[k]);
        __syncwarp();
        // ---- B: gather the 16 n0 of this lane's parity, radix-16, then the radix-2 across the lane pair
#pragma unroll
        for (int m = 0; m < 16; ++m) e[m] = exB[2 * m];
        Dft<16>::run(e);                                    // even lane: E[j]; odd lane: O[(j + 8) mod 16]
        float2 zl[8], zu[8];                                // bins k0 = 8 par + i  /  16 + 8 par + i
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float2 recv;
            recv.x = __shfl_xor_sync(0xffffffffu, e[8 + k].x, 1);
            recv.y = __shfl_xor_sync(0xffffffffu, e[8 + k].y, 1);
            // even lane: own = E[k], recv = O[k];  odd lane: own = O[8 + k], recv = E[8 + k]
            const float2 E = par ? recv : e[k];
            const float2 O = par ? e[k] : recv;
            const float2 T = cmul(O, rc[k]);                // W32^{k + 8 par} O
            zl[k] = E + T;
            zu[k] = E - T;
        }
        __syncwarp();
        // ---- split: publish the upper set, finish the pairs (k, M - k) of the lower set
#pragma unroll
        for (int i = 0; i < 8; ++i) zW[32 * i] = zu[i];
        __syncwarp();
        float2 b[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) b[i] = zR[32 * (7 - i)];
        if (k1 == 0) b[0] = par ? zu[0] : zl[0];            // k = 128 pairs with Z[384] (own upper 0); k = 0 with itself
        constexpr bool BANDS = OUT == OUT_BANDS_LOG || OUT == OUT_BANDS_LOW;
        if constexpr (BANDS) {
            // Fused band epilogue: scale |X| joins the eight band means of
            // examples/graph/utils.rs:85-116 (log set [0,4) [4,8) [8,16) ... [256,512)) or examples/pong/utils.rs:96-114
            // (eight 3-bin bands from bin 20); 40 bytes leave the SM per frame instead of 2052.
            // This lane's bins: k = k1 + 128 par + 16 i and M - k.  par = 1: k in [128, 256) = log band 6 for every i;
            // par = 0: i = 0 -> bands 0..2 by k1, i = 1 -> band 3, i = 2, 3 -> band 4, i >= 4 -> band 5; M - k is band 7.
            constexpr float ysc = WIN ? 1.0f : 0.5f;
            float part[4] = {0.f, 0.f, 0.f, 0.f};            // i = 0 | i = 1 | i = 2, 3 | i >= 4
            float hi = 0.0f;                                 // bins M - k: log band 7
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (OUT == OUT_BANDS_LOW && (i == 0 || i > 2)) continue;       // bins 20..43 only: i = 1, 2 of the par = 0 lanes
                const float2 A = zl[i], B = b[i];
                float2 S = make_float2(A.x + B.x, A.y - B.y);
                float2 D = make_float2(A.x - B.x, A.y + B.y);
                float2 Q = cmul(mul_w64(D, i), pw_lane);
                const float2 X1 = S + Q;
                part[i == 0 ? 0 : i == 1 ? 1 : i < 4 ? 2 : 3] += ysc * fast_sqrt(X1.x * X1.x + X1.y * X1.y);
                if constexpr (OUT == OUT_BANDS_LOG) {
                    const float2 X2 = make_float2(S.x - Q.x, Q.y - S.y);
                    const float y2 = ysc * fast_sqrt(X2.x * X2.x + X2.y * X2.y);
                    hi += (i == 0 && l == 0) ? 0.0f : y2;                      // k = 0 pairs with bin 512: not in a band
                }
            }
            // Warp reductions.  Every band's contributors are known lanes, so instead of eight full butterflies
            // (40 shuffles) the sums ride on lane parity and on runs of adjacent lane pairs; lane 0 collects.
            constexpr unsigned FULL = 0xffffffffu;
            float acc[8];
            if constexpr (OUT == OUT_BANDS_LOG) {
                if (l == 0) hi += 2.0f * ysc * fast_sqrt(zu[0].x * zu[0].x + zu[0].y * zu[0].y);   // bin 256
#pragma unroll
                for (int sft = 16; sft > 0; sft >>= 1) hi += __shfl_xor_sync(FULL, hi, sft);
                // odd lanes (bins 128..255): everything is band 6; even lanes: i >= 4 is band 5 -- one value per parity
                float v65 = par ? part[0] + part[1] + part[2] + part[3] : part[3];
                // even lanes: i = 2, 3 is band 4, i = 1 band 3: the band-3 term moves to the odd neighbour first
                const float p1 = __shfl_xor_sync(FULL, part[1], 1);
                float v43 = par ? p1 : part[2];
#pragma unroll
                for (int sft = 16; sft > 1; sft >>= 1) {
                    v65 += __shfl_xor_sync(FULL, v65, sft);
                    v43 += __shfl_xor_sync(FULL, v43, sft);
                }
                // i = 0 of the even lanes: bins k1 = l / 2 -> bands 0 (lanes 0..7), 1 (8..15), 2 (16..31)
                float v0 = par ? 0.0f : part[0];
                v0 += __shfl_xor_sync(FULL, v0, 2);
                v0 += __shfl_xor_sync(FULL, v0, 4);
                acc[0] = v0;                                                   // lane 0's own group
                acc[1] = __shfl_sync(FULL, v0, 8);
                acc[2] = __shfl_sync(FULL, v0, 16) + __shfl_sync(FULL, v0, 24);
                acc[3] = __shfl_sync(FULL, v43, 1);
                acc[4] = v43;
                acc[5] = v65;
                acc[6] = __shfl_sync(FULL, v65, 1);
                acc[7] = hi;
            } else {
                // bins 20..43 sit in the even lanes: i = 1 -> bin 16 + k1 (k1 >= 4), i = 2 -> bin 32 + k1 (k1 < 12); a band
                // is three consecutive k1 = three consecutive even lanes, summed by two shuffles from the lanes above
                float u = part[1], v = part[2];
                u += __shfl_down_sync(FULL, u, 2) + __shfl_down_sync(FULL, u, 4);
                v += __shfl_down_sync(FULL, v, 2) + __shfl_down_sync(FULL, v, 4);
#pragma unroll
                for (int bnd = 0; bnd < 4; ++bnd) {
                    acc[bnd] = __shfl_sync(FULL, u, 8 + 6 * bnd);              // k1 = 4 + 3 b: bins 20 + 3 b ..
                    acc[4 + bnd] = __shfl_sync(FULL, v, 6 * bnd);              // k1 = 3 b:     bins 32 + 3 b ..
                }
            }
            if (l == 0) {
                constexpr float inv_log[8] = {1.f / 4, 1.f / 4, 1.f / 8, 1.f / 16, 1.f / 32, 1.f / 64, 1.f / 128, 1.f / 256};
                float *row = reinterpret_cast<float *>(p.out) + frame * p.pitch;
                float mx = 0.0f;
                int idx = 0;
#pragma unroll
                for (int bnd = 0; bnd < 8; ++bnd) {
                    float m = acc[bnd] * p.band_scale * (OUT == OUT_BANDS_LOG ? inv_log[bnd] : (1.0f / 3.0f));
                    row[bnd] = m;
                    if (m > mx) { mx = m; idx = bnd; }                          // examples/pong/main.rs:200-207
                }
                row[8] = (float)idx;
                row[9] = mx;
            }
        } else {
            constexpr int EB = OUT == OUT_COMPLEX ? 8 : 4;
            char *o_lo = reinterpret_cast<char *>(p.out) + (frame * p.pitch + (k1 + 128 * par)) * EB;
            char *o_hi = reinterpret_cast<char *>(p.out) + (frame * p.pitch + (M - k1 - 128 * par)) * EB;
            if constexpr (BULK) {
                // the staging buffer is the warp's own and is never touched by the exchanges of the next frame
                if (l == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous frame has left
                __syncwarp();
                o_lo = reinterpret_cast<char *>(stage + (k1 + 128 * par));
                o_hi = reinterpret_cast<char *>(stage + (M - k1 - 128 * par));
            }
            // mirrors (FULL): bin k also lands at n - k; k = 0 (lane 0, i = 0, lower store) and k = 512 (its upper
            // store) have none
            char *m_lo = reinterpret_cast<char *>(p.out) + (frame * p.pitch + (2 * M - k1 - 128 * par)) * EB;
            char *m_hi = reinterpret_cast<char *>(p.out) + (frame * p.pitch + (M + k1 + 128 * par)) * EB;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float2 A = zl[i], B = b[i];
                float2 S = make_float2(A.x + B.x, A.y - B.y);
                float2 D = make_float2(A.x - B.x, A.y + B.y);
                float2 Q = cmul(mul_w64(D, i), pw_lane);        // -i W1024^{k1 + 128 par + 16 i}
                if constexpr (FULL) {
                    const bool edge = i == 0 && l == 0;         // bins 0 and 512: no mirror
                    emit_pair<OUT, WIN>(o_lo + 16 * i * EB, edge ? nullptr : m_lo - 16 * i * EB, S + Q);
                    emit_pair<OUT, WIN>(o_hi - 16 * i * EB, edge ? nullptr : m_hi + 16 * i * EB,
                                        make_float2(S.x - Q.x, Q.y - S.y));
                } else {
                    emit_at<OUT, WIN>(o_lo + 16 * i * EB, S + Q);
                    emit_at<OUT, WIN>(o_hi - 16 * i * EB, make_float2(S.x - Q.x, Q.y - S.y));
                }
            }
            if (l == 0) {                                                                                      // bin 256
                const float2 X256 = make_float2(2.0f * zu[0].x, -2.0f * zu[0].y);
                if constexpr (FULL) emit_pair<OUT, WIN>(o_lo + (M / 2) * EB, m_lo - (M / 2) * EB, X256);
                else emit_at<OUT, WIN>(o_lo + (M / 2) * EB, X256);
            }
            if constexpr (BULK) {
                __syncwarp();
                if (l == 0) {
                    float *row = reinterpret_cast<float *>(p.out) + frame * p.pitch;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 2048;" ::"l"(row),
                                 "r"(k1b_smem_u32(stage)) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    row[M] = stage[M];                             // bin 512: the 513th value of the row
                }
            }
        }
        __syncwarp();
    }
    if constexpr (BULK) {
        if (l == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

}  // namespace kopek
