// K1 -- batched real-input FFT with fused window and magnitude/dB epilogue (sm_100a).
//
// Replaces, for many frames at once, the caller blocks of the reference that
// widen f32 samples, call kopek::fft::fft (src/fft.rs:6-41) and take bin
// magnitudes (examples/beep/view.rs:140-158, examples/graph/utils.rs:47-63,
// examples/pong/utils.rs:36-47).  One HBM pass in (f32 samples), one out.
//
// Algorithm per frame of N real samples (M = N/2):
//   z[m] = w[2m] x[2m] + i w[2m+1] x[2m+1]           (window fused into the load)
//   Z    = FFT_M(z)   Stockham autosort, radix 4/8/16 butterflies in registers,
//                     stages exchanged through XOR-swizzled shared memory
//   X[k], X[M-k] from Z[k], conj Z[M-k]               (real-input split post-pass)
//   epilogue: complex | |X| | |X|^2 | 20 log10 |X|    written straight to HBM
// Each frame is owned by T = M/16 threads holding 16 complex values each.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace kopek {

enum : int { OUT_COMPLEX = 0, OUT_MAG = 1, OUT_POWER = 2, OUT_DB = 3, OUT_BANDS_LOG = 4, OUT_BANDS_LOW = 5 };

// Complex add / subtract.  With -DKOPEK_F32X2 the device side issues ONE packed instruction (add/sub.rn.f32x2, SASS
// FADD2) on the (re, im) register pair: same results bit for bit (two independent IEEE additions), same FP32 pipe
// time (tools/micro/fadd2.cu: 128 adds/clk/SM either way), half the issue slots.
#if defined(KOPEK_F32X2) && defined(__CUDA_ARCH__)
__device__ __forceinline__ float2 operator+(float2 a, float2 b) {
    float2 r;
    asm("{\n\t.reg .b64 pa, pb, pr;\n\tmov.b64 pa, {%2, %3};\n\tmov.b64 pb, {%4, %5};\n\tadd.rn.f32x2 pr, pa, pb;\n\t"
        "mov.b64 {%0, %1}, pr;\n\t}" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 operator-(float2 a, float2 b) {
    float2 r;
    asm("{\n\t.reg .b64 pa, pb, pr;\n\tmov.b64 pa, {%2, %3};\n\tmov.b64 pb, {%4, %5};\n\tsub.rn.f32x2 pr, pa, pb;\n\t"
        "mov.b64 {%0, %1}, pr;\n\t}" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
#else
__host__ __device__ __forceinline__ float2 operator+(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ float2 operator-(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#endif
__host__ __device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float fast_sqrt(float v) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float fast_log2(float v) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

// ---- radix butterflies: in place, natural-order output --------------------
__host__ __device__ __forceinline__ void dft4(float2 &v0, float2 &v1, float2 &v2, float2 &v3) {
    float2 s0 = v0 + v2, d0 = v0 - v2, s1 = v1 + v3, d1 = v1 - v3;
    v0 = s0 + s1;
    v2 = s0 - s1;
    v1 = make_float2(d0.x + d1.y, d0.y - d1.x);  // d0 - i d1
    v3 = make_float2(d0.x - d1.y, d0.y + d1.x);  // d0 + i d1
}

template <int R> struct Dft;
template <> struct Dft<2> {
    __host__ __device__ __forceinline__ static void run(float2 *v) {
        float2 a = v[0] + v[1], b = v[0] - v[1];
        v[0] = a; v[1] = b;
    }
};
template <> struct Dft<4> {
    __host__ __device__ __forceinline__ static void run(float2 *v) { dft4(v[0], v[1], v[2], v[3]); }
};
template <> struct Dft<8> {
    __host__ __device__ __forceinline__ static void run(float2 *v) {
        constexpr float h = 0.70710678118654752440f;
        float2 a0 = v[0] + v[4], a1 = v[1] + v[5], a2 = v[2] + v[6], a3 = v[3] + v[7];
        float2 b0 = v[0] - v[4], t1 = v[1] - v[5], t2 = v[2] - v[6], t3 = v[3] - v[7];
        float2 b1 = make_float2((t1.x + t1.y) * h, (t1.y - t1.x) * h);   // * W8^1
        float2 b2 = make_float2(t2.y, -t2.x);                            // * W8^2 = -i
        float2 b3 = make_float2((t3.y - t3.x) * h, -(t3.x + t3.y) * h);  // * W8^3
        dft4(a0, a1, a2, a3);
        dft4(b0, b1, b2, b3);
        v[0] = a0; v[1] = b0; v[2] = a1; v[3] = b1; v[4] = a2; v[5] = b2; v[6] = a3; v[7] = b3;
    }
};
// Second level of the radix-16 with the inter-level twiddles FOLDED into its additions (FMA form).
//   out = DFT4(a0, w1 a1, w2 a2, w3 a3),  w_t = W16^{t k1}
// The eighth-turn twiddles are h (p.x, p.y) with p a sum/difference of the components, so "a0 +- h p" is one FFMA per
// component instead of FMUL + FADD; the sixteenth-turn twiddles are c (x + t y, y - t x) with t = tan(pi/8), and the
// common factor c rides on the FFMAs of the output sums (both odd inputs of a row share |c|).  16 instructions fewer
// per radix-16 than multiply-then-butterfly (144 instead of 160), same operation count per output, one rounding fewer
// on the folded products.
#ifndef KOPEK_DFT16_PLAIN
template <int K1> __host__ __device__ __forceinline__ void dft16_row(float2 &v0, float2 &v1, float2 &v2, float2 &v3) {
    constexpr float h = 0.70710678118654752440f, c = 0.92387953251128675613f, t = 0.41421356237309504880f;
    if constexpr (K1 == 0) {
        dft4(v0, v1, v2, v3);
    } else if constexpr (K1 == 2) {
        // w1 = W16^2 = h (1 - i), w2 = -i, w3 = W16^6 = -h (1 + i)
        const float2 p1 = make_float2(v1.x + v1.y, v1.y - v1.x);          // a1 = h p1
        const float2 p3 = make_float2(v3.y - v3.x, -(v3.x + v3.y));       // a3 = h p3
        const float2 a2 = make_float2(v2.y, -v2.x);
        const float2 s0 = v0 + a2, d0 = v0 - a2;
        const float2 u = p1 + p3, w = p1 - p3;                            // s1 = h u, d1 = h w
        v0 = make_float2(fmaf(h, u.x, s0.x), fmaf(h, u.y, s0.y));
        v2 = make_float2(fmaf(-h, u.x, s0.x), fmaf(-h, u.y, s0.y));
        v1 = make_float2(fmaf(h, w.y, d0.x), fmaf(-h, w.x, d0.y));        // d0 - i d1
        v3 = make_float2(fmaf(-h, w.y, d0.x), fmaf(h, w.x, d0.y));        // d0 + i d1
    } else {
        // K1 = 1: w1 = W16^1 = (c, -s), w2 = W16^2, w3 = W16^3 = (s, -c)
        // K1 = 3: w1 = W16^3,           w2 = W16^6, w3 = W16^9 = -(c, -s)
        float2 q1, q3, p2;
        if constexpr (K1 == 1) {
            q1 = make_float2(fmaf(t, v1.y, v1.x), fmaf(-t, v1.x, v1.y));   // a1 = c q1
            q3 = make_float2(fmaf(t, v3.x, v3.y), fmaf(t, v3.y, -v3.x));   // a3 = c q3
            p2 = make_float2(v2.x + v2.y, v2.y - v2.x);                    // a2 = h p2
        } else {
            q1 = make_float2(fmaf(t, v1.x, v1.y), fmaf(t, v1.y, -v1.x));
            q3 = make_float2(fmaf(-t, v3.y, -v3.x), fmaf(t, v3.x, -v3.y)); // -(x + t y, y - t x)
            p2 = make_float2(v2.y - v2.x, -(v2.x + v2.y));
        }
        const float2 s0 = make_float2(fmaf(h, p2.x, v0.x), fmaf(h, p2.y, v0.y));
        const float2 d0 = make_float2(fmaf(-h, p2.x, v0.x), fmaf(-h, p2.y, v0.y));
        const float2 u = q1 + q3, w = q1 - q3;                             // s1 = c u, d1 = c w
        v0 = make_float2(fmaf(c, u.x, s0.x), fmaf(c, u.y, s0.y));
        v2 = make_float2(fmaf(-c, u.x, s0.x), fmaf(-c, u.y, s0.y));
        v1 = make_float2(fmaf(c, w.y, d0.x), fmaf(-c, w.x, d0.y));
        v3 = make_float2(fmaf(-c, w.y, d0.x), fmaf(c, w.x, d0.y));
    }
}
template <> struct Dft<16> {
    __host__ __device__ __forceinline__ static void run(float2 *v) {
#pragma unroll
        for (int t0 = 0; t0 < 4; ++t0) dft4(v[t0], v[t0 + 4], v[t0 + 8], v[t0 + 12]);
        // row k1 holds v[4 k1 + t0], still to be multiplied by W16^(t0 k1): folded into the row's butterfly
        dft16_row<0>(v[0], v[1], v[2], v[3]);
        dft16_row<1>(v[4], v[5], v[6], v[7]);
        dft16_row<2>(v[8], v[9], v[10], v[11]);
        dft16_row<3>(v[12], v[13], v[14], v[15]);
        // u[k1 + 4 k0] = v[4 k1 + k0]: transpose the 4x4 register tile
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = a + 1; b < 4; ++b) {
                float2 t = v[4 * a + b];
                v[4 * a + b] = v[4 * b + a];
                v[4 * b + a] = t;
            }
    }
};
#else
template <> struct Dft<16> {
    __host__ __device__ __forceinline__ static void run(float2 *v) {
        constexpr float h = 0.70710678118654752440f, c = 0.92387953251128675613f, s = 0.38268343236508977173f;
#pragma unroll
        for (int t0 = 0; t0 < 4; ++t0) dft4(v[t0], v[t0 + 4], v[t0 + 8], v[t0 + 12]);
        // v[t0 + 4 k1] *= W16^(t0 k1)
        v[5] = cmul(v[5], make_float2(c, -s));                        // W16^1
        v[9] = make_float2((v[9].x + v[9].y) * h, (v[9].y - v[9].x) * h);  // W16^2
        v[13] = cmul(v[13], make_float2(s, -c));                      // W16^3
        v[6] = make_float2((v[6].x + v[6].y) * h, (v[6].y - v[6].x) * h);  // W16^2
        v[10] = make_float2(v[10].y, -v[10].x);                       // W16^4 = -i
        v[14] = make_float2((v[14].y - v[14].x) * h, -(v[14].x + v[14].y) * h);  // W16^6
        v[7] = cmul(v[7], make_float2(s, -c));                        // W16^3
        v[11] = make_float2((v[11].y - v[11].x) * h, -(v[11].x + v[11].y) * h);  // W16^6
        v[15] = cmul(v[15], make_float2(-c, s));                      // W16^9
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
        // u[k1 + 4 k0] = v[4 k1 + k0]: transpose the 4x4 register tile
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = a + 1; b < 4; ++b) {
                float2 t = v[4 * a + b];
                v[4 * a + b] = v[4 * b + a];
                v[4 * b + a] = t;
            }
    }
};
#endif

// ---- radix 64 in registers: n = a + 4 b, k = kb + 16 ka:  X[k] = sum_a W4^{a ka} W64^{a kb} sum_b v[a + 4 b] W16^{b kb}
// W64^k = exp(-2 pi i k / 64); called with compile-time k after unrolling, so the values become immediates
__host__ __device__ __forceinline__ float2 k1_w64(int k) {
    constexpr float re[64] = {1.0f, 0.9951847266721969f, 0.9807852804032304f, 0.9569403357322088f, 0.9238795325112867f, 0.881921264348355f, 0.8314696123025452f, 0.773010453362737f, 0.7071067811865476f, 0.6343932841636455f, 0.5555702330196023f, 0.4713967368259978f, 0.38268343236508984f, 0.29028467725446233f, 0.19509032201612833f, 0.09801714032956077f, 0.0f, -0.09801714032956065f, -0.1950903220161282f, -0.29028467725446216f, -0.3826834323650897f, -0.4713967368259977f, -0.555570233019602f, -0.6343932841636454f, -0.7071067811865475f, -0.773010453362737f, -0.8314696123025453f, -0.8819212643483549f, -0.9238795325112867f, -0.9569403357322088f, -0.9807852804032304f, -0.9951847266721968f, -1.0f, -0.9951847266721969f, -0.9807852804032304f, -0.9569403357322089f, -0.9238795325112868f, -0.881921264348355f, -0.8314696123025455f, -0.7730104533627371f, -0.7071067811865477f, -0.6343932841636459f, -0.5555702330196022f, -0.47139673682599786f, -0.38268343236509034f, -0.29028467725446244f, -0.19509032201612866f, -0.09801714032956045f, 0.0f, 0.09801714032956009f, 0.1950903220161283f, 0.29028467725446205f, 0.38268343236509f, 0.4713967368259976f, 0.5555702330196018f, 0.6343932841636456f, 0.7071067811865474f, 0.7730104533627367f, 0.8314696123025452f, 0.8819212643483548f, 0.9238795325112865f, 0.9569403357322088f, 0.9807852804032303f, 0.9951847266721969f};
    constexpr float im[64] = {0.0f, -0.0980171403295606f, -0.19509032201612825f, -0.29028467725446233f, -0.3826834323650898f, -0.47139673682599764f, -0.5555702330196022f, -0.6343932841636455f, -0.7071067811865475f, -0.773010453362737f, -0.8314696123025452f, -0.8819212643483549f, -0.9238795325112867f, -0.9569403357322089f, -0.9807852804032304f, -0.9951847266721968f, -1.0f, -0.9951847266721969f, -0.9807852804032304f, -0.9569403357322089f, -0.9238795325112867f, -0.881921264348355f, -0.8314696123025455f, -0.7730104533627371f, -0.7071067811865476f, -0.6343932841636455f, -0.5555702330196022f, -0.47139673682599786f, -0.3826834323650899f, -0.2902846772544624f, -0.1950903220161286f, -0.09801714032956083f, 0.0f, 0.09801714032956059f, 0.19509032201612836f, 0.2902846772544621f, 0.38268343236508967f, 0.47139673682599764f, 0.555570233019602f, 0.6343932841636453f, 0.7071067811865475f, 0.7730104533627367f, 0.8314696123025452f, 0.8819212643483549f, 0.9238795325112865f, 0.9569403357322088f, 0.9807852804032303f, 0.9951847266721969f, 1.0f, 0.9951847266721969f, 0.9807852804032304f, 0.9569403357322089f, 0.9238795325112866f, 0.881921264348355f, 0.8314696123025455f, 0.7730104533627369f, 0.7071067811865477f, 0.6343932841636459f, 0.5555702330196022f, 0.4713967368259979f, 0.3826834323650904f, 0.2902846772544625f, 0.19509032201612872f, 0.0980171403295605f};
    return make_float2(re[k], im[k]);
}
template <> struct Dft<64> {
    __host__ __device__ __forceinline__ static void run(float2 *v) {
        float2 u[4][16];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
#pragma unroll
            for (int b = 0; b < 16; ++b) u[a][b] = v[a + 4 * b];
            Dft<16>::run(u[a]);
            if (a > 0) {
#pragma unroll
                for (int kb = 1; kb < 16; ++kb)
                    u[a][kb] = cmul(u[a][kb], k1_w64(a * kb));
            }
        }
#pragma unroll
        for (int kb = 0; kb < 16; ++kb) {
            dft4(u[0][kb], u[1][kb], u[2][kb], u[3][kb]);
#pragma unroll
            for (int ka = 0; ka < 4; ++ka) v[kb + 16 * ka] = u[ka][kb];
        }
    }
};

// ---- per-size configuration -------------------------------------------------
// register prefetch of the next frame group, per size (overridable for tuning runs: -DK1_PREF_ALL=0/1)
#ifdef K1_PREF_ALL
#define K1_PREF_256 K1_PREF_ALL
#define K1_PREF_512 K1_PREF_ALL
#define K1_PREF_1024 K1_PREF_ALL
#define K1_PREF_2048 K1_PREF_ALL
#define K1_PREF_4096 K1_PREF_ALL
#else   // measured on B200 (profiles/r1_configs.md): +1..4 points at every size
#define K1_PREF_256 true
#define K1_PREF_512 true
#define K1_PREF_1024 true
#define K1_PREF_2048 true
#define K1_PREF_4096 true
#endif
template <int N_> struct K1Cfg;
// radices, XOR swizzle q ^ ((q >> SWZ_SHIFT) & SWZ_MASK) and per-frame padding (complex slots) of the exchange
// buffer -- chosen with tools/bank_check.py (wavefronts within 1.08-1.13x of the conflict-free minimum)
template <> struct K1Cfg<256>  { static constexpr int S = 2, R0 = 8,  R1 = 16, R2 = 1;  static constexpr int SWZ_SHIFT = 3, SWZ_MASK = 7,  PAD = 8; static constexpr bool PREF = K1_PREF_256; };
template <> struct K1Cfg<512>  { static constexpr int S = 2, R0 = 16, R1 = 16, R2 = 1;  static constexpr int SWZ_SHIFT = 4, SWZ_MASK = 15, PAD = 0; static constexpr bool PREF = K1_PREF_512; };
template <> struct K1Cfg<1024> { static constexpr int S = 3, R0 = 8,  R1 = 8,  R2 = 8;  static constexpr int SWZ_SHIFT = 3, SWZ_MASK = 15, PAD = 0; static constexpr bool PREF = K1_PREF_1024; };
template <> struct K1Cfg<2048> { static constexpr int S = 3, R0 = 4,  R1 = 16, R2 = 16; static constexpr int SWZ_SHIFT = 4, SWZ_MASK = 15, PAD = 0; static constexpr bool PREF = K1_PREF_2048; };
template <> struct K1Cfg<4096> { static constexpr int S = 3, R0 = 8,  R1 = 16, R2 = 16; static constexpr int SWZ_SHIFT = 4, SWZ_MASK = 15, PAD = 0; static constexpr bool PREF = K1_PREF_4096; };

// Host: fill the M-entry stage-twiddle table of frame size N_ (stage 1 block, then stage 2 block).
template <int N_> inline void k1_build_stage_twiddles(float2 *out) {
    using C = K1Cfg<N_>;
    constexpr int M = N_ / 2;
    const double tau = 6.283185307179586476925286766559;
    for (int i = 0; i < M; ++i) out[i] = make_float2(1.0f, 0.0f);
    int off = 0;
    const int rad[2] = {C::R1, C::R2}, ns[2] = {C::R0, C::R0 * C::R1};
    for (int s = 0; s < 2; ++s) {
        if (rad[s] <= 1) break;
        for (int t = 1; t < rad[s]; ++t)
            for (int kk = 0; kk < ns[s]; ++kk) {
                double th = -tau * (double)(kk * t) / (double)(ns[s] * rad[s]);
                out[off + (t - 1) * ns[s] + kk] = make_float2((float)cos(th), (float)sin(th));
            }
        off += (rad[s] - 1) * ns[s];
    }
}

template <int N_> struct K1Geom {
    using C = K1Cfg<N_>;
    static constexpr int N = N_, M = N_ / 2, R = 16, T = M / R;
    static constexpr int BLOCK = T >= 128 ? T : 128;       // threads per CTA
    static constexpr int FPB = BLOCK / T;                  // frames per CTA per iteration
    static constexpr int PW = M / 2 + 2;                   // post twiddles (padded to even)
    // shared memory (bytes): stage twiddles, post twiddles, window, exchange buffers
    static constexpr int FSTRIDE = M + C::PAD;             // exchange-buffer slots per frame
    static constexpr size_t SMEM = sizeof(float2) * (M + PW + FPB * FSTRIDE) + sizeof(float) * N;
    __host__ __device__ static constexpr int swz(int q) { return q ^ ((q >> C::SWZ_SHIFT) & C::SWZ_MASK); }
};

struct K1Params {
    const float *x;        // samples
    void *out;             // spectra
    const float *win;      // N window coefficients, pre-scaled by 0.5 (nullptr: rectangular)
    const float2 *tw;      // per-stage t-major twiddle tables, M entries (k1_build_stage_twiddles)
    const float2 *pw;      // -i W_N^k = (-sin, -cos)(2 pi k / N), k <= M/2
    uint64_t n_frames;
    uint64_t hop;
    uint64_t pitch;        // output elements between rows
    int full;              // 1: also write the conjugate-symmetric bins M+1..N-1
    int vec;               // 1: 8-byte sample loads allowed (hop even, base 8-byte aligned)
};

// Epilogue of one bin.  `dst` points at the bin's element (float, or float2 for OUT_COMPLEX).
// X is 2x the spectrum when no window was applied (the 0.5 lives in the window table).
template <int OUT, bool WIN>
__device__ __forceinline__ void emit_at(void *dst, float2 X) {
    if constexpr (OUT == OUT_COMPLEX) {
        constexpr float sc = WIN ? 1.0f : 0.5f;
        *reinterpret_cast<float2 *>(dst) = make_float2(X.x * sc, X.y * sc);
    } else {
        float p = X.x * X.x + X.y * X.y;
        float r;
        if constexpr (OUT == OUT_MAG) r = WIN ? fast_sqrt(p) : 0.5f * fast_sqrt(p);
        else if constexpr (OUT == OUT_POWER) r = WIN ? p : 0.25f * p;
        else r = fmaf(fast_log2(p), 3.01029995663981195f, WIN ? 0.0f : -6.02059991327962390f);
        *reinterpret_cast<float *>(dst) = r;
    }
}
// Same value stored twice: at `dst` and (unless nullptr) at `mirror`, the bin n - k of a full-length row -- the
// conjugate for OUT_COMPLEX, the same magnitude / power / dB otherwise.
template <int OUT, bool WIN>
__device__ __forceinline__ void emit_pair(void *dst, void *mirror, float2 X) {
    if constexpr (OUT == OUT_COMPLEX) {
        constexpr float sc = WIN ? 1.0f : 0.5f;
        const float2 v = make_float2(X.x * sc, X.y * sc);
        *reinterpret_cast<float2 *>(dst) = v;
        if (mirror) *reinterpret_cast<float2 *>(mirror) = make_float2(v.x, -v.y);
    } else {
        float p = X.x * X.x + X.y * X.y;
        float r;
        if constexpr (OUT == OUT_MAG) r = WIN ? fast_sqrt(p) : 0.5f * fast_sqrt(p);
        else if constexpr (OUT == OUT_POWER) r = WIN ? p : 0.25f * p;
        else r = fmaf(fast_log2(p), 3.01029995663981195f, WIN ? 0.0f : -6.02059991327962390f);
        *reinterpret_cast<float *>(dst) = r;
        if (mirror) *reinterpret_cast<float *>(mirror) = r;
    }
}
template <int OUT, bool WIN>
__device__ __forceinline__ void emit(void *out, uint64_t idx, float2 X) {
    constexpr int EB = OUT == OUT_COMPLEX ? 8 : 4;
    emit_at<OUT, WIN>(reinterpret_cast<char *>(out) + idx * EB, X);
}

template <int RAD, int NS, int M, int T, class G>
__device__ __forceinline__ void stage_write(float2 *buf, const float2 *d, int tid) {
    constexpr int NB = 16 / RAD;
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        int j = tid + T * i;
        int kk = j & (NS - 1);
        int base = (j - kk) * RAD + kk;
#pragma unroll
        for (int k = 0; k < RAD; ++k) buf[G::swz(base + k * NS)] = d[i * RAD + k];
    }
}

// tw is the stage's own table, t-major: tw[(t-1) * NS + kk] = W_M^{kk t M/(NS RAD)}, so that lanes with
// consecutive kk read consecutive words (no bank conflicts; v1 indexed one W_M^e table with stride kk*t).
template <int RAD, int NS, int M, int T, class G>
__device__ __forceinline__ void stage_read_twiddle(const float2 *buf, const float2 *tw, float2 *d, int tid) {
    constexpr int NB = 16 / RAD;
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        int j = tid + T * i;
        int kk = j & (NS - 1);
#pragma unroll
        for (int t = 0; t < RAD; ++t) {
            float2 v = buf[G::swz(j + t * (M / RAD))];
            if (t > 0) v = cmul(v, tw[(t - 1) * NS + kk]);
            d[i * RAD + t] = v;
        }
    }
}

template <int RAD> __device__ __forceinline__ void stage_dft(float2 *d) {
#pragma unroll
    for (int i = 0; i < 16 / RAD; ++i) Dft<RAD>::run(d + i * RAD);
}

template <int T> __device__ __forceinline__ void frame_sync() {
    if constexpr (T <= 32) __syncwarp();
    else __syncthreads();
}

// PREF: the samples of the CTA's NEXT frame group are loaded into registers at the top of the current one
// (32 registers; pays off where the register budget keeps the occupancy: see K1Cfg<N>::PREF).
template <int N_, int OUT, bool WIN>
__global__ void __launch_bounds__(K1Geom<N_>::BLOCK)
k1_stft_kernel(const K1Params p) {
    constexpr bool PREF = K1Cfg<N_>::PREF;
    using G = K1Geom<N_>;
    using C = K1Cfg<N_>;
    constexpr int M = G::M, T = G::T, R0 = C::R0, R1 = C::R1, R2 = C::R2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *s_tw = reinterpret_cast<float2 *>(smem_raw);
    float2 *s_pw = s_tw + M;
    float2 *s_ex = s_pw + G::PW;
    float *s_win = reinterpret_cast<float *>(s_ex + G::FPB * G::FSTRIDE);

    for (int i = threadIdx.x; i < M; i += G::BLOCK) s_tw[i] = p.tw[i];
    for (int i = threadIdx.x; i <= M / 2; i += G::BLOCK) s_pw[i] = p.pw[i];
    if constexpr (WIN)
        for (int i = threadIdx.x; i < N_; i += G::BLOCK) s_win[i] = p.win[i];
    __syncthreads();

    const int tid = threadIdx.x % T;
    const int fslot = threadIdx.x / T;
    float2 *buf = s_ex + fslot * G::FSTRIDE;
    const uint64_t n_groups = (p.n_frames + G::FPB - 1) / G::FPB;

    // samples of frame group gg, slot fslot -> registers (d[i*R0 + t] = z[j + t M/R0], j = tid + T i)
    auto load_frame = [&](uint64_t gg, float2 *dst) {
        uint64_t fr = gg * G::FPB + fslot;
        if (fr >= p.n_frames) fr = p.n_frames - 1;
        const float *xf = p.x + fr * p.hop;
#pragma unroll
        for (int i = 0; i < 16 / R0; ++i) {
            int j = tid + T * i;
#pragma unroll
            for (int t = 0; t < R0; ++t) {
                int q = j + t * (M / R0);
                if (p.vec) dst[i * R0 + t] = __ldg(reinterpret_cast<const float2 *>(xf) + q);
                else dst[i * R0 + t] = make_float2(__ldg(xf + 2 * q), __ldg(xf + 2 * q + 1));
            }
        }
    };
    float2 nx[16];
    if constexpr (PREF) {
        if (blockIdx.x < n_groups) load_frame(blockIdx.x, nx);
    }

    for (uint64_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
        uint64_t frame = g * G::FPB + fslot;
        const bool valid = frame < p.n_frames;
        if (!valid) frame = p.n_frames - 1;

        float2 d[16];
        if constexpr (PREF) {
#pragma unroll
            for (int i = 0; i < 16; ++i) d[i] = nx[i];
            if (g + gridDim.x < n_groups) load_frame(g + gridDim.x, nx);
        } else {
            load_frame(g, d);
        }
        // ---- stage 0: window, butterflies over stride M/R0
#pragma unroll
        for (int i = 0; i < 16 / R0; ++i) {
            int j = tid + T * i;
#pragma unroll
            for (int t = 0; t < R0; ++t) {
                int q = j + t * (M / R0);
                float2 v = d[i * R0 + t];
                if constexpr (WIN) {
                    float2 w = reinterpret_cast<const float2 *>(s_win)[q];
                    v.x *= w.x;
                    v.y *= w.y;
                }
                d[i * R0 + t] = v;
            }
        }
        stage_dft<R0>(d);
        stage_write<R0, 1, M, T, G>(buf, d, tid);
        frame_sync<T>();
        // ---- stage 1
        stage_read_twiddle<R1, R0, M, T, G>(buf, s_tw, d, tid);
        stage_dft<R1>(d);
        frame_sync<T>();
        stage_write<R1, R0, M, T, G>(buf, d, tid);
        frame_sync<T>();
        // ---- stage 2
        if constexpr (R2 > 1) {
            stage_read_twiddle<R2, R0 * R1, M, T, G>(buf, s_tw + (R1 - 1) * R0, d, tid);
            stage_dft<R2>(d);
            frame_sync<T>();
            stage_write<R2, R0 * R1, M, T, G>(buf, d, tid);
            frame_sync<T>();
        }
        // ---- real-input split + epilogue.  Pair (k, M-k):
        //   S = Z[k] + conj Z[M-k], D = Z[k] - conj Z[M-k], Q = (-i W_N^k) D
        //   2 X[k] = S + Q,  2 X[M-k] = conj(S - Q)      (k = 0 gives X[0] and X[M])
        const uint64_t row = frame * p.pitch;
#pragma unroll
        for (int i = 0; i <= 8; ++i) {
            int k = tid + T * i;
            if (i == 8 && tid != 0) break;       // the lone k = M/2 pair
            float2 a = buf[G::swz(k)];
            float2 b = buf[G::swz((M - k) & (M - 1))];
            float2 S = make_float2(a.x + b.x, a.y - b.y);
            float2 D = make_float2(a.x - b.x, a.y + b.y);
            float2 Q = cmul(D, s_pw[k]);
            float2 Xk = S + Q;
            float2 Xm = make_float2(S.x - Q.x, -(S.y - Q.y));
            if (valid) {
                emit<OUT, WIN>(p.out, row + k, Xk);
                if (i < 8) emit<OUT, WIN>(p.out, row + (M - k), Xm);
                if (p.full) {
                    // X[N-k] = conj X[k]
                    if (k > 0) emit<OUT, WIN>(p.out, row + (N_ - k), make_float2(Xk.x, -Xk.y));
                    if (i < 8 && k > 0) emit<OUT, WIN>(p.out, row + (M + k), make_float2(Xm.x, -Xm.y));
                }
            }
        }
        frame_sync<T>();
    }
}

}  // namespace kopek
