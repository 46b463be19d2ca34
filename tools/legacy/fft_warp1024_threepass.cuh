// K1W -- warp-per-frame 1024-point real-input FFT + fused window / magnitude epilogue (sm_100a).
//
// The headline kernel of the path (BASELINE.json metric: batched 1024-pt FFT+mag).
// Same mathematics as fft_batched.cuh (K1) -- window, 512-point complex FFT of the
// packed frame, real-input split, |X| / dB -- but laid out so that ONE WARP owns one
// frame end to end and the shared-memory data pipe (the measured bottleneck of K1,
// profiles/r1_v1_k1_full.md: 92 % busy) moves the minimum number of 128-byte wavefronts:
//
//   n = 64 n2 + 8 n1 + n0,  k = k2 + 8 k1 + 64 k0          (three radix-8 passes, DIF order)
//   A: lane (n1, n0 pair) transforms over n2, multiplies by W512^{(8 n1 + n0) k2}
//   B: lane (k2, n0 pair) transforms over n1 IN PLACE (reads and writes its own slots),
//      multiplies by W64^{n0 k1}
//   C: lane (k2, k1 pair) transforms over n0 -> Z[k2 + 8 k1 + 64 k0] in registers
//   split: the partner Z[M-k] of every bin a lane owns lives in ONE float4 of lane 35-l,
//      so only the upper half (k0 >= 4) is published and each lane finishes 8 (k, M-k) pairs.
//
// All exchanges are 128-bit, conflict free by construction (one spare float4 per eight:
// slot L sits at L + L/8), and every address is "per-lane base + immediate".
// tools/k1w_model.py is the numpy model of exactly these address functions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fft_batched.cuh"

namespace kopek {

struct K1WParams {
    const float *x;          // samples (16-byte aligned, hop % 4 == 0)
    void *out;               // spectra
    const float4 *tables;    // K1W_TABLE_F4 float4: window | TPP | TQ | PW (see k1w_build_tables)
    uint64_t n_frames;
    uint64_t hop;
    uint64_t pitch;
    float band_scale;        // OUT_BANDS_*: display scale applied to the band means
    uint32_t pcm_shift;      // PCM input (fft_warp1024b.cuh): 16 = channel 0 (low half of a stereo frame), 0 = channel 1
};

constexpr int K1W_WIN_F4 = 256;            // 0.5 * w[4i .. 4i+3]
constexpr int K1W_TPP_F4 = 7 * 32;         // [k2-1][lane]  (W256^{lane k2}, W256^{lane k2} W512^{k2})
constexpr int K1W_TQ_F4 = 7 * 4;           // [k1-1][b]     (W32^{b k1},    W32^{b k1} W64^{k1})
constexpr int K1W_PW_F4 = 4 * 32;          // [k0][lane]    (-i W1024^{k}, -i W1024^{k+8}), k = k2 + 16 c + 64 k0
constexpr int K1W_TABLE_F4 = K1W_WIN_F4 + K1W_TPP_F4 + K1W_TQ_F4 + K1W_PW_F4;
constexpr int K1W_EX_F4 = 288;             // per-warp exchange buffer: 256 slots + 32 spare

constexpr int K1W_IN_F4 = 256;             // per-warp TMA landing buffer: one frame, linear

constexpr int K1W_STAGE_F4 = 132;          // per-warp output staging of the bulk-store epilogue (516 floats + pad)

template <int WARPS, bool TMA = false, bool BULK = false> struct K1WGeom {
    static constexpr int BLOCK = WARPS * 32;
    static constexpr int WARP_F4 = K1W_EX_F4 + (TMA ? K1W_IN_F4 : 0) + (BULK ? K1W_STAGE_F4 : 0);
    // tables | per-warp (exchange [+ landing]) | per-warp mbarrier
    static constexpr size_t SMEM = sizeof(float4) * (K1W_TABLE_F4 + WARPS * WARP_F4) + (TMA ? 8 * WARPS : 0);
};

// ---- TMA (bulk async copy) + mbarrier primitives: one frame = one 4096-byte cp.async.bulk ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "KOPEK_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra KOPEK_DONE;\n\t"
        "bra KOPEK_WAIT;\n\t"
        "KOPEK_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// multiply by a compile-time constant (c, s): 2 FMUL + 2 FFMA with immediates
__device__ __forceinline__ float2 cmul_const(float2 a, float c, float s) {
    return make_float2(fmaf(-a.y, s, a.x * c), fmaf(a.y, c, a.x * s));
}
// W512^k and W64^k, k = 1..7 (the odd-element factors of passes A and B)
__device__ __forceinline__ float2 mul_w512(float2 a, int k) {
    constexpr float c[8] = {1.0f, 0.99992470183914454092f, 0.99969881869620422012f, 0.99932238458834950090f,
                            0.99879545620517239271f, 0.99811811290014920526f, 0.99729045667869021614f,
                            0.99631261218277801050f};
    constexpr float s[8] = {0.0f, -0.01227153828571992608f, -0.02454122852291228803f, -0.03680722294135883231f,
                            -0.04906767432741801426f, -0.06132073630220857779f, -0.07356456359966742631f,
                            -0.08579731234443989385f};
    return cmul_const(a, c[k], s[k]);
}
__device__ __forceinline__ float2 mul_w64(float2 a, int k) {
    constexpr float c[8] = {1.0f, 0.99518472667219688624f, 0.98078528040323044913f, 0.95694033573220886494f,
                            0.92387953251128675613f, 0.88192126434835502971f, 0.83146961230254523708f,
                            0.77301045336273696081f};
    constexpr float s[8] = {0.0f, -0.09801714032956060199f, -0.19509032201612826785f, -0.29028467725446236764f,
                            -0.38268343236508977173f, -0.47139673682599764856f, -0.55557023301960222474f,
                            -0.63439328416364549822f};
    return cmul_const(a, c[k], s[k]);
}

// ---- fused band epilogue (OUT_BANDS_LOG / OUT_BANDS_LOW) -----------------------------------------
// acc[b] += y for the band b that bin k falls in; only bands LO..HI are possible for this call site.
template <int LO, int HI> __device__ __forceinline__ void band_add(float (&acc)[8], int band, float y) {
#pragma unroll
    for (int b = LO; b <= HI; ++b) acc[b] += band == b ? y : 0.0f;
}
// log-spaced bands [0,4) [4,8) [8,16) ... [256,512): index = max(0, floor(log2 k) - 1)
__device__ __forceinline__ int band_log_of(int k) { return max(0, 30 - __clz(k)); }
// eight 3-bin bands from bin 20: index (k - 20) / 3 for 20 <= k < 44, else none (-1)
__device__ __forceinline__ int band_low_of(int k) {
    return (k >= 20 && k < 44) ? ((k - 20) * 11) >> 5 : -1;      // x*11>>5 == x/3 for 0 <= x < 24
}

// W1024^{8 f + 64 k0} (index 2 k0 + f): the compile-time part of the split twiddle -i W1024^{kbase + 8 f + 64 k0};
// the lane part -i W1024^{kbase} is one register pair.  Saves the 16 wavefronts of a post-twiddle table per frame.
__device__ __forceinline__ float2 mul_w1024_off(float2 a, int idx) {
    constexpr float c[8] = {1.0f, 0.99879545620517240501f, 0.92387953251128673848f, 0.90398929312344333820f,
                            0.70710678118654757274f, 0.67155895484701833009f, 0.38268343236508983729f,
                            0.33688985339222005111f};
    constexpr float s[8] = {0.0f, -0.04906767432741801493f, -0.38268343236508978178f, -0.42755509343028208491f,
                            -0.70710678118654746172f, -0.74095112535495910588f, -0.92387953251128673848f,
                            -0.94154406518302080631f};
    return idx == 0 ? a : cmul_const(a, c[idx], s[idx]);
}

// TWREG: 0 = pass-A/B twiddles from shared-memory tables; 2 = all four sets live in registers (loop invariant per
//        lane, 56 registers); 1 = only the even-element sets do (28 registers), the odd element's twiddle is the
//        even one times a compile-time rotation (W512^{k2}, W64^{k1}: +56 FP instructions per frame);
// WINREG: so does the lane's slice of the window.
// TMA: the NEXT frame of the warp is fetched by one cp.async.bulk into the warp's landing buffer while
//      the current one is transformed (the buffer is free as soon as pass A has read it).
// LOAD: 0 = plain LDG.128 at the top of the frame, 1 = TMA prefetch (above), 2 = LDG.128 of the next frame
//       into registers at the top of the current one (no shared-memory landing traffic, 32 more registers).
// BULK (f32 outputs, rows 16-byte aligned: pitch % 4 == 0): the frame's 513 values are staged in the warp's exchange
//       buffer and leave as ONE cp.async.bulk store of 2048 bytes (+ bin 512) instead of 17 32-bit stores per lane.
//       On local HBM the direct stores are as fast; over NVLink (peer-direct spectrogram, kopek_b200/sharded.py) the
//       bulk store is what reaches link bandwidth.
template <int OUT, bool WIN, int WARPS, int MINB, int TWREG, bool WINREG, int LOAD, bool BULK = false>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k1w_1024_kernel(const K1WParams p) {
    constexpr int M = 512;
    static_assert(!BULK || (OUT == OUT_MAG || OUT == OUT_POWER || OUT == OUT_DB), "bulk store: f32 spectra only");
    constexpr bool TMA = LOAD == 1;
    extern __shared__ __align__(16) unsigned char k1w_smem[];
    float4 *s_win = reinterpret_cast<float4 *>(k1w_smem);
    float4 *s_tpp = s_win + K1W_WIN_F4;
    float4 *s_tq = s_tpp + K1W_TPP_F4;
    float4 *s_pw = s_tq + K1W_TQ_F4;
    float4 *s_ex = s_pw + K1W_PW_F4;

    for (int i = threadIdx.x + (WIN ? 0 : K1W_WIN_F4); i < K1W_TABLE_F4; i += WARPS * 32) s_win[i] = p.tables[i];
    __syncthreads();

    const int l = threadIdx.x & 31;
    float4 *ex = s_ex + (threadIdx.x >> 5) * K1WGeom<WARPS, TMA, BULK>::WARP_F4;
    const float4 *in = ex + K1W_EX_F4 + l;                   // TMA landing buffer, + 32 n2
    uint64_t *bar = reinterpret_cast<uint64_t *>(s_ex + WARPS * K1WGeom<WARPS, TMA, BULK>::WARP_F4) + (threadIdx.x >> 5);
    uint32_t phase = 0;
    float4 *exA = ex + l + (l >> 3);                         // + 36 k2
    float4 *exB = ex + 36 * (l >> 2) + (l & 3);              // + 4 n1 + (n1 >> 1)
    float4 *exC = ex + 36 * (l >> 2) + 9 * (l & 3);          // + 4 f + s
    float4 *zW = ex + ((l + 4) & 31);                        // + 32 (k0 - 4)
    const float4 *zR = ex + (((l >= 4 ? 35 - l : 3 - l) + 4) & 31);   // + 32 (3 - k0)
    const float4 *tpp = s_tpp + l;                           // + 32 (k2 - 1)
    const float4 *tq = s_tq + (l & 3);                       // + 4 (k1 - 1)
    const float4 *pwt = s_pw + l;                            // + 32 k0
    const float4 *wt = s_win + l;                            // + 32 n2
    const int kbase = (l >> 2) + 16 * (l & 3);               // k = kbase + 8 f + 64 k0
    const float2 pw_lane = make_float2(pwt[0].x, pwt[0].y);  // -i W1024^{kbase}
    float2 rp[8], rp1[8], rq[8], rq1[8];                     // TWREG: W256^{l k2}, W512^{(2l+1) k2}, W32^{b k1}, W64^{(2b+1) k1}
    float4 rw[8];                                            // WINREG: window slice
    if constexpr (TWREG) {
#pragma unroll
        for (int k = 1; k < 8; ++k) {
            float4 t = tpp[32 * (k - 1)], u = tq[4 * (k - 1)];
            rp[k] = make_float2(t.x, t.y);
            rp1[k] = make_float2(t.z, t.w);
            rq[k] = make_float2(u.x, u.y);
            rq1[k] = make_float2(u.z, u.w);
        }
    }
    if constexpr (WIN && WINREG) {
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) rw[n2] = wt[32 * n2];
    }

    const uint64_t total_warps = (uint64_t)gridDim.x * WARPS;
    const uint64_t first = (uint64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
    if constexpr (TMA) {
        if (l == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (first < p.n_frames) {
                mbar_expect_tx(bar, 4096);
                tma_load_1d(ex + K1W_EX_F4, p.x + first * p.hop, 4096, bar);
            }
        }
        __syncwarp();
    }
    float4 vn[8];                                            // LOAD == 2: the next frame, in flight
    if constexpr (LOAD == 2) {
        if (first < p.n_frames) {
            const float4 *xf = reinterpret_cast<const float4 *>(p.x + first * p.hop) + l;
#pragma unroll
            for (int n2 = 0; n2 < 8; ++n2) vn[n2] = ldg_stream(xf + 32 * n2);
        }
    }
    for (uint64_t frame = first; frame < p.n_frames; frame += total_warps) {
        // ---- A: load + window, transform over n2, twiddle, publish
        float4 v[8];
        if constexpr (LOAD == 2) {
#pragma unroll
            for (int n2 = 0; n2 < 8; ++n2) v[n2] = vn[n2];
            if (frame + total_warps < p.n_frames) {
                const float4 *xf = reinterpret_cast<const float4 *>(p.x + (frame + total_warps) * p.hop) + l;
#pragma unroll
                for (int n2 = 0; n2 < 8; ++n2) vn[n2] = ldg_stream(xf + 32 * n2);
            }
        } else if constexpr (TMA) {
            mbar_wait(bar, phase);
            phase ^= 1;
#pragma unroll
            for (int n2 = 0; n2 < 8; ++n2) v[n2] = in[32 * n2];
            __syncwarp();                                    // every lane has its samples: the buffer is free
            if (l == 0 && frame + total_warps < p.n_frames) {
                mbar_expect_tx(bar, 4096);
                tma_load_1d(ex + K1W_EX_F4, p.x + (frame + total_warps) * p.hop, 4096, bar);
            }
        } else {
            const float4 *xf = reinterpret_cast<const float4 *>(p.x + frame * p.hop) + l;
#pragma unroll
            for (int n2 = 0; n2 < 8; ++n2) v[n2] = ldg_stream(xf + 32 * n2);
        }
        float2 d0[8], d1[8];
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) {
            if constexpr (WIN) {
                float4 w;
                if constexpr (WINREG) w = rw[n2];
                else w = wt[32 * n2];
                d0[n2] = make_float2(v[n2].x * w.x, v[n2].y * w.y);
                d1[n2] = make_float2(v[n2].z * w.z, v[n2].w * w.w);
            } else {
                d0[n2] = make_float2(v[n2].x, v[n2].y);
                d1[n2] = make_float2(v[n2].z, v[n2].w);
            }
        }
        Dft<8>::run(d0);
        Dft<8>::run(d1);
        exA[0] = make_float4(d0[0].x, d0[0].y, d1[0].x, d1[0].y);
#pragma unroll
        for (int k2 = 1; k2 < 8; ++k2) {
            float2 a, b;
            if constexpr (TWREG) {
                a = cmul(d0[k2], rp[k2]);
                if constexpr (TWREG == 2) b = cmul(d1[k2], rp1[k2]);
                else b = cmul(mul_w512(d1[k2], k2), rp[k2]);
            } else {
                float4 t = tpp[32 * (k2 - 1)];
                a = cmul(d0[k2], make_float2(t.x, t.y));
                b = cmul(d1[k2], make_float2(t.z, t.w));
            }
            exA[36 * k2] = make_float4(a.x, a.y, b.x, b.y);
        }
        __syncwarp();
        // ---- B: in place over n1
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
            float4 t = exB[4 * n1 + (n1 >> 1)];
            d0[n1] = make_float2(t.x, t.y);
            d1[n1] = make_float2(t.z, t.w);
        }
        Dft<8>::run(d0);
        Dft<8>::run(d1);
        exB[0] = make_float4(d0[0].x, d0[0].y, d1[0].x, d1[0].y);
#pragma unroll
        for (int k1 = 1; k1 < 8; ++k1) {
            float2 a, b;
            if constexpr (TWREG) {
                a = cmul(d0[k1], rq[k1]);
                if constexpr (TWREG == 2) b = cmul(d1[k1], rq1[k1]);
                else b = cmul(mul_w64(d1[k1], k1), rq[k1]);
            } else {
                float4 t = tq[4 * (k1 - 1)];
                a = cmul(d0[k1], make_float2(t.x, t.y));
                b = cmul(d1[k1], make_float2(t.z, t.w));
            }
            exB[4 * k1 + (k1 >> 1)] = make_float4(a.x, a.y, b.x, b.y);
        }
        __syncwarp();
        // ---- C: over n0 -> Z[kbase + 8 f + 64 k0] (d0: f = 0, d1: f = 1)
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            float4 t = exC[s];
            d0[2 * s] = make_float2(t.x, t.y);
            d0[2 * s + 1] = make_float2(t.z, t.w);
            float4 u = exC[4 + s];
            d1[2 * s] = make_float2(u.x, u.y);
            d1[2 * s + 1] = make_float2(u.z, u.w);
        }
        Dft<8>::run(d0);
        Dft<8>::run(d1);
        __syncwarp();
        // ---- split: publish the upper half, finish the pairs (k, M-k) for k0 < 4
#pragma unroll
        for (int k0 = 4; k0 < 8; ++k0) zW[32 * (k0 - 4)] = make_float4(d0[k0].x, d0[k0].y, d1[k0].x, d1[k0].y);
        __syncwarp();
        float2 b0[4], b1[4];      // partners Z[M-k] of the f = 0 / f = 1 bins
#pragma unroll
        for (int k0 = 0; k0 < 4; ++k0) {
            float4 t = zR[32 * (3 - k0)];
            b0[k0] = make_float2(t.z, t.w);
            b1[k0] = l >= 4 ? make_float2(t.x, t.y) : make_float2(t.z, t.w);
        }
        if (l < 4) {              // k2 = 0 lanes: the f = 0 partner sits in a different slot
            const float4 *z2 = ex + ((((-l) & 3) + 4) & 31);
#pragma unroll
            for (int k0 = 0; k0 < 4; ++k0) {
                if (l == 0 && k0 == 0) {
                    b0[0] = d0[0];                              // k = 0: Z[M] == Z[0]
                } else {
                    float4 t = z2[32 * ((l == 0 ? 8 - k0 : 7 - k0) - 4)];
                    b0[k0] = make_float2(t.x, t.y);
                }
            }
        }
        // bins k = kbase + (8 f + 64 k0) ascend from o_lo, bins M - k descend from o_hi: immediates only
        constexpr bool BANDS = OUT == OUT_BANDS_LOG || OUT == OUT_BANDS_LOW;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        constexpr int EB = OUT == OUT_COMPLEX ? 8 : 4;
        char *o_lo = reinterpret_cast<char *>(p.out) + (frame * p.pitch + kbase) * EB;
        char *o_hi = reinterpret_cast<char *>(p.out) + (frame * p.pitch + (M - kbase)) * EB;
        // BULK: the warp's own staging buffer (never aliased by the exchange passes of the next frame, which start
        // while the TMA engine may still be reading this frame's row)
        float *stage = reinterpret_cast<float *>(ex + K1W_EX_F4 + (TMA ? K1W_IN_F4 : 0));
        if constexpr (BULK) {
            if (l == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous frame has left
            __syncwarp();
            o_lo = reinterpret_cast<char *>(stage + kbase);
            o_hi = reinterpret_cast<char *>(stage + (M - kbase));
        }
#pragma unroll
        for (int k0 = 0; k0 < 4; ++k0) {
#pragma unroll
            for (int f = 0; f < 2; ++f) {
                const int koff = (8 * f + 64 * k0) * EB;
                const float2 a = f ? d1[k0] : d0[k0];
                const float2 b = f ? b1[k0] : b0[k0];
                float2 S = make_float2(a.x + b.x, a.y - b.y);
                float2 D = make_float2(a.x - b.x, a.y + b.y);
                float2 Q = cmul(mul_w1024_off(D, 2 * k0 + f), pw_lane);
                if constexpr (BANDS) {
                    // |X[k]| joins its band; |X[M-k]| (bins 256..511) is band 7 of the log set
                    const float2 X1 = S + Q;
                    const float y1 = (WIN ? 1.0f : 0.5f) * fast_sqrt(X1.x * X1.x + X1.y * X1.y);
                    const int k = kbase + 8 * f + 64 * k0;
                    if constexpr (OUT == OUT_BANDS_LOG) {
                        if (k0 == 0 && f == 0) band_add<0, 4>(acc, band_log_of(k), y1);
                        else if (k0 == 0) band_add<2, 5>(acc, band_log_of(k), y1);
                        else if (k0 == 1 && f == 0) acc[5] += y1;
                        else if (k0 == 1) band_add<5, 6>(acc, band_log_of(k), y1);
                        else if (k0 == 2 || f == 0) acc[6] += y1;
                        else band_add<6, 7>(acc, band_log_of(k), y1);
                        const float2 X2 = make_float2(S.x - Q.x, Q.y - S.y);
                        const float y2 = (WIN ? 1.0f : 0.5f) * fast_sqrt(X2.x * X2.x + X2.y * X2.y);
                        if (k0 == 0 && f == 0) acc[7] += k > 0 ? y2 : 0.0f;      // k = 0 pairs with bin 512: not in a band
                        else acc[7] += y2;
                    } else {
                        if (k0 == 0) band_add<0, 7>(acc, band_low_of(k), y1);
                    }
                } else {
                    emit_at<OUT, WIN>(o_lo + koff, S + Q);
                    emit_at<OUT, WIN>(o_hi - koff, make_float2(S.x - Q.x, Q.y - S.y));
                }
            }
        }
        if constexpr (BANDS) {
            if constexpr (OUT == OUT_BANDS_LOG) {   // bin 256 = conj Z[256] (lane 0) is in band 7
                if (l == 0) acc[7] += (WIN ? 2.0f : 1.0f) * fast_sqrt(d0[4].x * d0[4].x + d0[4].y * d0[4].y);
            }
#pragma unroll
            for (int b = 0; b < 8; ++b) {
#pragma unroll
                for (int sft = 16; sft > 0; sft >>= 1) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], sft);
            }
            if (l == 0) {
                constexpr float inv_log[8] = {1.f / 4, 1.f / 4, 1.f / 8, 1.f / 16, 1.f / 32, 1.f / 64, 1.f / 128, 1.f / 256};
                float *row = reinterpret_cast<float *>(p.out) + frame * p.pitch;
                float mx = 0.0f;
                int idx = 0;
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    float m = acc[b] * p.band_scale * (OUT == OUT_BANDS_LOG ? inv_log[b] : (1.0f / 3.0f));
                    row[b] = m;
                    if (m > mx) { mx = m; idx = b; }           // examples/pong/main.rs:200-207
                }
                row[8] = (float)idx;
                row[9] = mx;
            }
        } else {
            if (l == 0) emit_at<OUT, WIN>(o_lo + (M / 2) * EB, make_float2(2.0f * d0[4].x, -2.0f * d0[4].y));
            if constexpr (BULK) {
                __syncwarp();
                if (l == 0) {
                    float *row = reinterpret_cast<float *>(p.out) + frame * p.pitch;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 2048;" ::"l"(row),
                                 "r"(smem_u32(stage)) : "memory");
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
