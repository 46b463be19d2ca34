// K1M -- team-per-frame real-input FFT + fused window / magnitude epilogue for N = 2048 and 4096 (sm_100a).
//
// BASELINE.json configs[1] (Hann 2048, hop 512, dB) and configs[4] (Hann 4096, hop 1024) run on this kernel.
// It carries the two-exchange structure of the 1024-point kernel (fft_warp1024b.cuh) to T = 16 R lanes per frame
// (R = 4: two warps, N = 2048;  R = 8: four warps, N = 4096;  R = 2 is the 1024-point layout, kept for testing):
//
//   M = N/2 = 16 T,  n = T n1 + n0,  k = k1 + 16 k0,  k0 = j + 16 q
//   A: lane t = n0 loads z[T n1 + t] (n1 = 0..15), radix-16 over n1, multiplies by W_M^{n0 k1}, publishes (k1, n0)
//   B: the T-point transform over n0 of butterfly k1 belongs to the R adjacent lanes t = R k1 + r: lane r gathers
//      n0 = R m + r and runs a radix-16 over m.  The remaining radix-R across the lane group is an ALL-TO-ALL by
//      shfl.idx: this lane finishes the 16/R values j = c_r + d (c_r = (16/R) r) for every q and needs exactly
//      those j from each other lane.  Pass A publishes lane n0's values times W_R^{(n0 % R)(n0 / R)}, which ROTATES
//      the radix-16 output of lane r by c_r (register i holds E_r[(i + c_r) mod 16]); reading from lane
//      (r + x) mod R then makes the wanted j sit in register (16 - (16/R) x + d) mod 16 of EVERY sender: no selects.
//      Receiver: twiddle W_T^{r' j}, R-point DFT over x, phase W_R^{r q} (the rotation x = r' - r).
//      (R - 1) 16/R complex shuffles per lane: 12 (R = 4) or 14 (R = 8), against 32/48 for radix-2 lane stages.
//   split: every finished Z[k1 + 16 j'] (j' = j + 16 q) is written to shared memory at (T + 1) k1 + j'; after the
//      second barrier lane t finishes the pairs (k, M - k) for k = t + T i (i = 0..7): consecutive lanes hold
//      consecutive bins, so every global store is a full 128-byte run (the first version stored 16-byte runs from
//      the (k1, r) layout and was bound by the store path: 60 % of the HBM roofline at N = 4096).
// Shared memory: element (k1, n0) at 64-bit position (T + R) k1 + n0, Z at (T + 1) k1 + j': every access is conflict
// free (tools/k1m_model.py is the numpy model of the pass A/B address functions, checked against numpy.fft.rfft).
// Two CTA barriers per frame; every loop-invariant table lives in registers; the next frame is prefetched into
// registers at the top of the current one.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fft_warp512.cuh"

namespace kopek {


template <int R> struct K1MGeom {
    static constexpr int T = 16 * R, QD = 16 / R, M = 16 * T, N = 32 * T, S = T + R, S2 = T + 1;
    static constexpr int WIN_F2 = M;             // [T n1 + t]   (0.5 w[2i], 0.5 w[2i+1])
    static constexpr int RP_F2 = 16 * T;         // [k1][t]      W_M^{t k1} W_R^{(t % R)(t / R)}
    static constexpr int TW_F2 = 16 * T;         // [x QD + d][t] W_T^{((r + x) % R)(c_r + d)}
    static constexpr int PH_F2 = (R / 2) * T;    // [q - 1][t]   W_R^{r q}, q = 1 .. R/2 - 1 (row R/2 - 1 unused)
    static constexpr int PW_F2 = T;              // [t]          -i W_N^{t}
    static constexpr int CS_F2 = 5 * T;          // cosine-sum windows: [c][t] (P_c, Q_c), c = 0, 1;  [2][t] (a/2, -);
                                                 // [3][t], [4][t]: the table values of n1 = 0 and n1 = 15 (frame edges)
    static constexpr int TABLE_F2 = WIN_F2 + RP_F2 + TW_F2 + PH_F2 + PW_F2 + CS_F2;
    static constexpr int EX_F2 = 16 * S;         // pass A -> B exchange
    static constexpr int ZX_F2 = 16 * S2;        // the finished Z, row k1 (S2 = T + 1 positions), column j'
    // (The partner loads of the row-0 lanes put a third request on one bank pair: 16 extra wavefronts per warp and
    // frame.  A shifted copy of row 0 in a 17th row removes them -- and was 2.5-3 points SLOWER, measured: the first
    // warp, which writes the copy, becomes the straggler every other warp of the team waits for at the barrier.)
    static constexpr size_t SMEM = sizeof(float2) * (EX_F2 + ZX_F2);
    static constexpr size_t SMEM_WIN = sizeof(float2) * M;     // WINK == 3: the window table, N floats
};

// multiply by W256^m, m a compile-time index < 64 (the compile-time part of the split twiddle)
__device__ __forceinline__ float2 mul_w256(float2 v, int m) {
    constexpr float c[64] = {
        1.00000000000000000000f, 0.99969881869620424997f, 0.99879545620517240501f, 0.99729045667869020697f,
        0.99518472667219692873f, 0.99247953459870996706f, 0.98917650996478101444f, 0.98527764238894122162f,
        0.98078528040323043058f, 0.97570213003852857003f, 0.97003125319454397424f, 0.96377606579543984022f,
        0.95694033573220882438f, 0.94952818059303667475f, 0.94154406518302080631f, 0.93299279883473895669f,
        0.92387953251128673848f, 0.91420975570353069095f, 0.90398929312344333820f, 0.89322430119551532446f,
        0.88192126434835504956f, 0.87008699110871146054f, 0.85772861000027211809f, 0.84485356524970711689f,
        0.83146961230254523567f, 0.81758481315158371139f, 0.80320753148064494287f, 0.78834642762660633863f,
        0.77301045336273699338f, 0.75720884650648456748f, 0.74095112535495910588f, 0.72424708295146700276f,
        0.70710678118654757274f, 0.68954054473706694051f, 0.67155895484701833009f, 0.65317284295377675551f,
        0.63439328416364548779f, 0.61523159058062681925f, 0.59569930449243346793f, 0.57580819141784533866f,
        0.55557023301960228867f, 0.53499761988709726435f, 0.51410274419322166128f, 0.49289819222978409341f,
        0.47139673682599780857f, 0.44961132965460659516f, 0.42755509343028219593f, 0.40524131400498986100f,
        0.38268343236508983729f, 0.35989503653498827740f, 0.33688985339222005111f, 0.31368174039889157312f,
        0.29028467725446233105f, 0.26671275747489842090f, 0.24298017990326398197f, 0.21910124015686976984f,
        0.19509032201612833135f, 0.17096188876030135595f, 0.14673047445536174793f, 0.12241067519921627893f,
        0.09801714032956077016f, 0.07356456359966745406f, 0.04906767432741812596f, 0.02454122852291226384f,
    };
    constexpr float s[64] = {
        -0.00000000000000000000f, -0.02454122852291228812f, -0.04906767432741801493f, -0.07356456359966742631f,
        -0.09801714032956060363f, -0.12241067519921619566f, -0.14673047445536174793f, -0.17096188876030121717f,
        -0.19509032201612824808f, -0.21910124015686979759f, -0.24298017990326387094f, -0.26671275747489836538f,
        -0.29028467725446233105f, -0.31368174039889151761f, -0.33688985339222005111f, -0.35989503653498811087f,
        -0.38268343236508978178f, -0.40524131400498986100f, -0.42755509343028208491f, -0.44961132965460653965f,
        -0.47139673682599764204f, -0.49289819222978403790f, -0.51410274419322166128f, -0.53499761988709715332f,
        -0.55557023301960217765f, -0.57580819141784533866f, -0.59569930449243335691f, -0.61523159058062681925f,
        -0.63439328416364548779f, -0.65317284295377675551f, -0.67155895484701833009f, -0.68954054473706682948f,
        -0.70710678118654746172f, -0.72424708295146689174f, -0.74095112535495910588f, -0.75720884650648445646f,
        -0.77301045336273699338f, -0.78834642762660622761f, -0.80320753148064483184f, -0.81758481315158371139f,
        -0.83146961230254523567f, -0.84485356524970700587f, -0.85772861000027211809f, -0.87008699110871134952f,
        -0.88192126434835493853f, -0.89322430119551532446f, -0.90398929312344333820f, -0.91420975570353069095f,
        -0.92387953251128673848f, -0.93299279883473884567f, -0.94154406518302080631f, -0.94952818059303667475f,
        -0.95694033573220893540f, -0.96377606579543984022f, -0.97003125319454397424f, -0.97570213003852857003f,
        -0.98078528040323043058f, -0.98527764238894122162f, -0.98917650996478101444f, -0.99247953459870996706f,
        -0.99518472667219681771f, -0.99729045667869020697f, -0.99879545620517240501f, -0.99969881869620424997f,
    };
    return m == 0 ? v : cmul_const(v, c[m], s[m]);
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {      // a * conj(b)
    return make_float2(fmaf(a.y, b.y, a.x * b.x), fmaf(a.y, b.x, -a.x * b.y));
}
// multiply by W_4^m = (-i)^m, m a compile-time constant: register renaming and sign flips only
__device__ __forceinline__ float2 rot_w4(float2 v, int m) {
    return m == 0 ? v : m == 1 ? make_float2(v.y, -v.x) : m == 2 ? make_float2(-v.x, -v.y) : make_float2(-v.y, v.x);
}
__device__ __forceinline__ float2 ldg_stream2(const float2 *p) {
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}

// one 128-byte line per lane into L2, no register destination: the DRAM latency of the frame the team will load
// into registers next-but-PF is paid here, so the register prefetch only has to cover the L2 latency
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// cos(n1 pi/8), sin(n1 pi/8): the compile-time part of a cosine-sum window (see WINK == 2 below)
__device__ __forceinline__ float cs16_cos(int n1) {
    constexpr float c[5] = {1.0f, 0.92387953251128673848f, 0.70710678118654757274f, 0.38268343236508983729f, 0.0f};
    const int m = n1 & 7;
    const float v = m <= 4 ? c[m] : -c[8 - m];
    return (n1 & 8) ? -v : v;
}
__device__ __forceinline__ float cs16_sin(int n1) { return cs16_cos((n1 + 12) & 15); }

// WINK: 0 = rectangular, 3 = window table in SHARED memory (any window, used for KOPEK_WINDOW_CUSTOM: the team's CTA
// holds the N coefficients as [j][t] float4 = (w(2j, c = 0), w(2j, 1), w(2j + 1, 0), w(2j + 1, 1)), so a lane reads
// its 32 coefficients with eight conflict-free 128-bit loads per frame and no window register is held; the round-1
// form, 32 window registers and 8 warps per SM, was 4 / 1.3 points slower at 2048 / 4096 points and is gone),
// 2 = cosine-sum window
// w[n] = a - b cos(2 pi n / N) (Hann, Hamming) evaluated on the fly: the samples of lane t sit at angles
// theta_{t,c} + n1 pi/8 (n = 2 (T n1 + t) + c), so w/2 = a/2 + P_c cos(n1 pi/8) + Q_c sin(n1 pi/8) with the lane
// constants P_c = -(b/2) cos theta, Q_c = (b/2) sin theta: 9 registers instead of 32, which is what lets three
// 128-lane teams (or six 64-lane teams) stay resident per SM, and it beats the shared-memory table by 3-4 points.  The sum has an absolute error of ~3e-8, which is a
// large RELATIVE error where the window itself is ~1e-7 (the first and last samples of a frame), so the two edge
// sixteenths (n1 = 0 and n1 = 15) keep their table values; everywhere else w/2 >= 0.019 and the error is < 2e-6.
// FULL: all N bins per row (the conjugate mirror X[N - k] stored beside the computed half, see fft_warp1024b.cuh);
// VEC == false: frames start at any 4-byte aligned sample (odd hop / unaligned base): two 32-bit loads per pair.
template <int R, int OUT, int WINK, int MINB, int PF, bool FULL = false, bool VEC = true>
__global__ void __launch_bounds__(16 * R, MINB)
k1m_kernel(const K1WParams p) {
    constexpr bool WIN = WINK != 0;
    using G = K1MGeom<R>;
    constexpr int T = G::T, QD = G::QD, M = G::M, S = G::S, S2 = G::S2;
    extern __shared__ __align__(16) unsigned char k1m_smem[];
    float2 *ex = reinterpret_cast<float2 *>(k1m_smem);
    float2 *zs = ex + G::EX_F2;                              // Z[k1 + 16 j'] at S2 k1 + j'
    float4 *s_win = reinterpret_cast<float4 *>(zs + G::ZX_F2);   // WINK == 3 only: [j][t]
    const int t = threadIdx.x, k1 = t / R, r = t % R;
    float2 *exA = ex + t;                                    // + S k1'
    const float2 *exB = ex + S * k1 + r;                     // + R m
    float2 *zW = zs + S2 * k1 + QD * r;                      // + 16 q + d
    const float2 *zA = zs + S2 * (t & 15) + (t >> 4);        // + R i          Z[t + T i]
    // Z[M - k]: row 16 - k1', column T - 1 - j'; the lanes of row 0 (k1' = 0) pair inside their own row, column T - j'
    const float2 *zB = zs + ((t & 15) ? S2 * (16 - (t & 15)) + (T - 1 - (t >> 4)) : T - (t >> 4));   // - R i

    const float2 *tab = reinterpret_cast<const float2 *>(p.tables);
    if constexpr (WINK == 3) {
        // (the exchange buffers hold 16 (T + R) + 16 (T + 1) float2, a multiple of two: s_win is 16-byte aligned)
        for (int i = threadIdx.x; i < 8 * T; i += T) {
            const int j = i / T, tt = i % T;
            const float2 a = tab[T * (2 * j) + tt], b = tab[T * (2 * j + 1) + tt];
            s_win[i] = make_float4(a.x, a.y, b.x, b.y);
        }
        if constexpr (T == 32) __syncwarp(); else __syncthreads();
    }
    float2 rp[16], tw[16], ph[R / 2];
    float2 cs0, cs1, cs_first, cs_last;
    float cs_a;
    if constexpr (WINK == 2) {
        const float2 *cs = tab + G::WIN_F2 + G::RP_F2 + G::TW_F2 + G::PH_F2 + G::PW_F2;
        cs0 = cs[t];
        cs1 = cs[T + t];
        cs_a = cs[2 * T + t].x;
        cs_first = cs[3 * T + t];
        cs_last = cs[4 * T + t];
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) rp[k] = tab[G::WIN_F2 + T * k + t];
#pragma unroll
    for (int k = 0; k < 16; ++k) tw[k] = tab[G::WIN_F2 + G::RP_F2 + T * k + t];
#pragma unroll
    for (int q = 1; q < R / 2; ++q) ph[q] = tab[G::WIN_F2 + G::RP_F2 + G::TW_F2 + T * (q - 1) + t];
    const float2 pw_lane = tab[G::WIN_F2 + G::RP_F2 + G::TW_F2 + G::PH_F2 + t];
    const float sgn = (r & 1) ? -1.0f : 1.0f;               // W_R^{r R/2}

    auto load_frame = [&](uint64_t fr, float2 *dst) {
        if constexpr (VEC) {
            const float2 *xf = reinterpret_cast<const float2 *>(p.x + fr * p.hop) + t;
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) dst[n1] = ldg_stream2(xf + T * n1);
        } else {
            const float *xs = p.x + fr * p.hop + 2 * t;
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) dst[n1] = make_float2(__ldg(xs + 2 * T * n1), __ldg(xs + 2 * T * n1 + 1));
        }
    };
    float2 vn[16];
    if (blockIdx.x < p.n_frames) load_frame(blockIdx.x, vn);
    // lane t owns line t of a frame (N floats = T lines); with hop < N only the lines holding new samples are fetched
    const bool pf_lane = 32 * (t + 1) + p.hop > (uint64_t)G::N;
    if constexpr (PF > 1) {
#pragma unroll
        for (int a = 1; a < PF; ++a)
            if (pf_lane && blockIdx.x + (uint64_t)a * gridDim.x < p.n_frames)
                prefetch_l2(p.x + (blockIdx.x + (uint64_t)a * gridDim.x) * p.hop + 32 * t);
    }

    for (uint64_t frame = blockIdx.x; frame < p.n_frames; frame += gridDim.x) {
        if constexpr (PF > 0) {
            const uint64_t fp = frame + (uint64_t)PF * gridDim.x;
            if (pf_lane && fp < p.n_frames) prefetch_l2(p.x + fp * p.hop + 32 * t);
        }
        float2 e[16];
        if constexpr (WINK == 3) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 w = s_win[T * j + t];
                e[2 * j] = make_float2(vn[2 * j].x * w.x, vn[2 * j].y * w.y);
                e[2 * j + 1] = make_float2(vn[2 * j + 1].x * w.z, vn[2 * j + 1].y * w.w);
            }
        }
        if constexpr (WINK == 2) {
            // the window values are loop invariant: without this the compiler hoists all 32 of them into registers
            asm volatile("" : "+f"(cs_a), "+f"(cs0.x), "+f"(cs0.y), "+f"(cs1.x), "+f"(cs1.y));
        }
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
            if constexpr (WINK == 3) {
                // done above
            } else if constexpr (WINK == 2) {
                const float c = cs16_cos(n1), sn = cs16_sin(n1);
                float w0, w1;
                if (n1 == 0) {
                    w0 = cs_first.x, w1 = cs_first.y;
                } else if (n1 == 15) {
                    w0 = cs_last.x, w1 = cs_last.y;
                } else if (n1 == 8) {                        // sin = 0, cos = -1
                    w0 = cs_a - cs0.x;
                    w1 = cs_a - cs1.x;
                } else if (n1 % 8 == 4) {                    // cos = 0, sin = +-1
                    w0 = cs_a + (n1 == 12 ? -cs0.y : cs0.y);
                    w1 = cs_a + (n1 == 12 ? -cs1.y : cs1.y);
                } else {
                    w0 = fmaf(cs0.y, sn, fmaf(cs0.x, c, cs_a));
                    w1 = fmaf(cs1.y, sn, fmaf(cs1.x, c, cs_a));
                }
                e[n1] = make_float2(vn[n1].x * w0, vn[n1].y * w1);
            } else {
                e[n1] = vn[n1];
            }
        }
        if (frame + gridDim.x < p.n_frames) load_frame(frame + gridDim.x, vn);

        // ---- A: radix-16 over n1, twiddle (pre-rotation folded in), publish (k1', n0 = t)
        Dft<16>::run(e);
#pragma unroll
        for (int k = 0; k < 16; ++k) exA[S * k] = cmul(e[k], rp[k]);
        if constexpr (T == 32) __syncwarp(); else __syncthreads();
        // ---- B: gather n0 = R m + r, radix-16 (register i = E_r[(i + c_r) mod 16]), all-to-all, radix-R
#pragma unroll
        for (int m = 0; m < 16; ++m) e[m] = exB[R * m];
        Dft<16>::run(e);
        // Z[k1 + 16 j'] (j' = j + 16 q) goes to position S2 k1 + j' as soon as it is finished
#pragma unroll
        for (int d = 0; d < QD; ++d) {
            float2 g[R];
            g[0] = cmul(e[d], tw[d]);
#pragma unroll
            for (int x = 1; x < R; ++x) {
                const int reg = (16 - QD * x + d) & 15;
                float2 recv;
                recv.x = __shfl_sync(0xffffffffu, e[reg].x, r + x, R);
                recv.y = __shfl_sync(0xffffffffu, e[reg].y, r + x, R);
                g[x] = cmul(recv, tw[x * QD + d]);
            }
            Dft<R>::run(g);
            if constexpr (R == 4) {
                // the phase W_4^{r q} is a quarter turn, and for the READER of this element (r, q) are compile-time
                // functions of its pair index i (below): stored unphased, rotated for free on the other side
#pragma unroll
                for (int q = 0; q < R; ++q) zW[16 * q + d] = g[q];
            } else {
                zW[d] = g[0];
#pragma unroll
                for (int q = 1; q < R / 2; ++q) zW[16 * q + d] = cmul(g[q], ph[q]);
                zW[16 * (R / 2) + d] = make_float2(g[R / 2].x * sgn, g[R / 2].y * sgn);
#pragma unroll
                for (int q = 1; q < R / 2; ++q) zW[16 * (R / 2 + q) + d] = cmul_conj(g[R / 2 + q], ph[R / 2 - q]);
            }
        }
        if constexpr (T == 32) __syncwarp(); else __syncthreads();
        // ---- split: lane t finishes the pairs (k, M - k), k = t + T i: consecutive lanes, consecutive bins
        float2 a[8], b[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            a[i] = zA[R * i];
            b[i] = zB[-R * i];
        }
        if constexpr (R == 4) {
            // writer of Z[t + 64 i]: (r, q) = (i % 4, i / 4); of its partner Z[M - k]: (3 - i % 4, 3 - i / 4).  Lane 0
            // alone pairs inside row 0 one column further (j' -> 64 - j'), which shifts its partners' (r, q).
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a[i] = rot_w4(a[i], ((i & 3) * (i >> 2)) & 3);
                b[i] = rot_w4(b[i], ((3 - (i & 3)) * (3 - (i >> 2))) & 3);
            }
            if (t == 0) {
#pragma unroll
                for (int i = 1; i < 8; ++i) b[i] = rot_w4(b[i], i < 4 ? 3 : 2);
            }
        }
        if (t == 0) b[0] = a[0];                             // k = 0 pairs with itself (Z[M] = Z[0])
        constexpr int EB = OUT == OUT_COMPLEX ? 8 : 4;
        char *o_lo = reinterpret_cast<char *>(p.out) + (frame * p.pitch + t) * EB;
        char *o_hi = reinterpret_cast<char *>(p.out) + (frame * p.pitch + (M - t)) * EB;
        // mirrors (FULL): bin k also lands at N - k; k = 0 and k = M (lane 0, i = 0) have none
        char *m_lo = reinterpret_cast<char *>(p.out) + (frame * p.pitch + (2 * M - t)) * EB;
        char *m_hi = reinterpret_cast<char *>(p.out) + (frame * p.pitch + (M + t)) * EB;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 A = a[i], B = b[i];
            float2 Sx = make_float2(A.x + B.x, A.y - B.y);
            float2 Dx = make_float2(A.x - B.x, A.y + B.y);
            float2 Qx = cmul(mul_w32(Dx, i), pw_lane);       // -i W_N^{t + T i} = (-i W_N^t) W32^i
            if constexpr (FULL) {
                const bool edge = i == 0 && t == 0;
                emit_pair<OUT, WIN>(o_lo + T * i * EB, edge ? nullptr : m_lo - T * i * EB, Sx + Qx);
                emit_pair<OUT, WIN>(o_hi - T * i * EB, edge ? nullptr : m_hi + T * i * EB,
                                    make_float2(Sx.x - Qx.x, Qx.y - Sx.y));
            } else {
                emit_at<OUT, WIN>(o_lo + T * i * EB, Sx + Qx);
                emit_at<OUT, WIN>(o_hi - T * i * EB, make_float2(Sx.x - Qx.x, Qx.y - Sx.y));
            }
        }
        if (t == 0) {                                        // bin N/4 = conj Z[M/2]
            const float2 zm = zs[T / 2];
            const float2 Xq = make_float2(2.0f * zm.x, -2.0f * zm.y);
            if constexpr (FULL) emit_pair<OUT, WIN>(o_lo + (M / 2) * EB, m_lo - (M / 2) * EB, Xq);
            else emit_at<OUT, WIN>(o_lo + (M / 2) * EB, Xq);
        }
    }
}

}  // namespace kopek
