// K2 -- exact f64 complex FFT, the device side of the kopek::fft::fft drop-in.
//
// src/fft.rs:6-41 is a recursive radix-2 decimation-in-time FFT: the level that
// combines two length-L/2 halves computes, for k < L/2 (src/fft.rs:23-27),
//     t = w_k * odd[k];  out[k] = even[k] + t;  out[k + L/2] = even[k] - t
// with w_k = exp(-i pi i'/n), i' = 2*step*k, step = n/L, built as
// (cos th, sin th), th = ((-PI) * i') / n.  The recursion only decides WHERE
// values live; the VALUES are those of the textbook iterative DIT on a
// bit-reversed input.  This kernel evaluates exactly those butterflies: same
// twiddles (table filled on the host with the same libm expression), same four
// products and two sums per complex multiply, no fused multiply-add
// (__dmul_rn/__dadd_rn are never contracted).  Result: bit-identical to the
// f64 recursion, for every n.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace kopek {

__device__ __forceinline__ double2 k2_butterfly_t(double2 w, double2 b) {
    // Complex Mul of num-complex: (w.re*b.re - w.im*b.im, w.re*b.im + w.im*b.re)
    return make_double2(__dsub_rn(__dmul_rn(w.x, b.x), __dmul_rn(w.y, b.y)),
                        __dadd_rn(__dmul_rn(w.x, b.y), __dmul_rn(w.y, b.x)));
}

__device__ __forceinline__ uint32_t k2_bitrev(uint32_t v, int bits) {
    return bits == 0 ? 0u : (__brev(v) >> (32 - bits));
}

// Levels L = 2 .. 2^chunk_log on a chunk of 2^chunk_log consecutive positions of
// the bit-reversed sequence, resident in shared memory.  grid = (chunks, batch).
// tw[j] = twiddle i' = 2j of the full length n = 2^logn  (j < n/2).
__global__ void k2_c2c_f64_chunk(const double2 *__restrict__ in, size_t n_in, size_t in_stride,
                                 double2 *__restrict__ out, size_t out_stride,
                                 const double2 *__restrict__ tw, int logn, int chunk_log) {
    extern __shared__ double2 k2_smem[];
    const uint32_t cn = 1u << chunk_log;
    const size_t base = (size_t)blockIdx.x << chunk_log;
    in += (size_t)blockIdx.y * in_stride;
    out += (size_t)blockIdx.y * out_stride;

    // Position base+p of the bit-reversed order holds input element bitrev(base+p) = (bitrev_chunk(p) << hi) + bitrev_hi(chunk).
    // The loop runs over the SOURCE index so that a whole transform in one chunk (n <= 4096: every UI-rate call) reads its
    // input in order -- full lines even when `in` is mapped host memory read over PCIe -- and scatters in shared memory.
    const int hi = logn - chunk_log;
    const size_t src0 = k2_bitrev(blockIdx.x, hi);
    for (uint32_t q = threadIdx.x; q < cn; q += blockDim.x) {
        size_t src = ((size_t)q << hi) + src0;
        k2_smem[k2_bitrev(q, chunk_log)] = src < n_in ? in[src] : make_double2(0.0, 0.0);   // zero pad, src/fft.rs:36
    }
    __syncthreads();
    for (int lv = 1; lv <= chunk_log; ++lv) {
        const uint32_t half = 1u << (lv - 1);
        const int tshift = logn - lv;                     // twiddle index j = k * n / L
        for (uint32_t b = threadIdx.x; b < cn / 2; b += blockDim.x) {
            uint32_t k = b & (half - 1);
            uint32_t i0 = ((b - k) << 1) + k, i1 = i0 + half;
            double2 t = k2_butterfly_t(tw[(size_t)k << tshift], k2_smem[i1]);
            double2 a = k2_smem[i0];
            k2_smem[i0] = make_double2(__dadd_rn(a.x, t.x), __dadd_rn(a.y, t.y));
            k2_smem[i1] = make_double2(__dsub_rn(a.x, t.x), __dsub_rn(a.y, t.y));
        }
        __syncthreads();
    }
    for (uint32_t p = threadIdx.x; p < cn; p += blockDim.x) out[base + p] = k2_smem[p];
}

// One level L = 2^lv > chunk in global memory, in place.  One butterfly per thread.
__global__ void k2_c2c_f64_level(double2 *__restrict__ data, const double2 *__restrict__ tw, int logn, int lv) {
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t half = (size_t)1 << (lv - 1);
    if (b >= ((size_t)1 << (logn - 1))) return;
    size_t k = b & (half - 1);
    size_t i0 = ((b - k) << 1) + k, i1 = i0 + half;
    double2 t = k2_butterfly_t(tw[k << (logn - lv)], data[i1]);
    double2 a = data[i0];
    data[i0] = make_double2(__dadd_rn(a.x, t.x), __dadd_rn(a.y, t.y));
    data[i1] = make_double2(__dsub_rn(a.x, t.x), __dsub_rn(a.y, t.y));
}

}  // namespace kopek
