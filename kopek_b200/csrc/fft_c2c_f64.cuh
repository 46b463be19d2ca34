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
//
// Latency is what this path is judged on (one 1024-point call per UI frame), so a thread carries 2^R elements
// through R consecutive levels in registers -- the same radix-2 butterflies on the same operands in the same
// order per element, only the trips through shared memory and the barriers between levels are shared: a
// 1024-point transform is 4 passes instead of 10.
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

// Shared-memory position of element e of a 2^L-point chunk (16-byte elements: a quarter warp is conflict free when
// its eight lanes differ in the low three bits).  The first term spreads the lanes of the first pass (lane = group
// of 2^R consecutive elements), the second the lanes of the bit-reversal scatter (lane = top bits of e).
__device__ __forceinline__ uint32_t k2_pos(uint32_t e, int L) {
    uint32_t p = e ^ ((e >> 3) & 7u);
    if (L >= 9) p ^= (e >> (L - 3)) & 7u;
    return p;
}

// Levels lv0 .. lv0+R-1 for the 2^R elements base + j h, h = 2^(lv0-1), of work item `item`.
// ld(e) / st(e, v) access element e of the sequence.  At level lv0+s (half H = h << s) element j with bit s clear is
// the upper input of the butterfly with k = k0 + (j mod 2^s) h, its partner is element j + 2^s; the twiddle is
// tw[k << (logn - lv)] like in the one-level form.
template <int R, class LD, class ST>
__device__ __forceinline__ void k2_levels(const double2 *__restrict__ tw, int logn, int lv0, size_t item, LD ld, ST st) {
    constexpr int E = 1 << R;
    const size_t h = (size_t)1 << (lv0 - 1);
    const size_t k0 = item & (h - 1);
    const size_t base = ((item - k0) << R) + k0;
    double2 w[E - 1];
#pragma unroll
    for (int s = 0; s < R; ++s)
#pragma unroll
        for (int m = 0; m < (1 << s); ++m) w[(1 << s) - 1 + m] = tw[(k0 + (size_t)m * h) << (logn - lv0 - s)];
    double2 v[E];
#pragma unroll
    for (int j = 0; j < E; ++j) v[j] = ld(base + (size_t)j * h);
#pragma unroll
    for (int s = 0; s < R; ++s)
#pragma unroll
        for (int j = 0; j < E; ++j)
            if (!(j & (1 << s))) {
                const double2 t = k2_butterfly_t(w[(1 << s) - 1 + (j & ((1 << s) - 1))], v[j + (1 << s)]);
                const double2 a = v[j];
                v[j] = make_double2(__dadd_rn(a.x, t.x), __dadd_rn(a.y, t.y));
                v[j + (1 << s)] = make_double2(__dsub_rn(a.x, t.x), __dsub_rn(a.y, t.y));
            }
#pragma unroll
    for (int j = 0; j < E; ++j) st(base + (size_t)j * h, v[j]);
}

// Levels 1 .. chunk_log on a chunk of 2^chunk_log consecutive positions of the bit-reversed sequence, resident in
// shared memory.  grid = (chunks, batch), max(32, chunk / 8) threads.  tw[j] = twiddle i' = 2j of the full length
// n = 2^logn (j < n/2).  `flag` (optional, single-CTA launches on the host-buffer path): once the spectrum has been
// stored, `token` is written there -- mapped host memory the caller polls instead of paying for a stream
// synchronisation.
__global__ void k2_c2c_f64_chunk(const double2 *__restrict__ in, size_t n_in, size_t in_stride,
                                 double2 *__restrict__ out, size_t out_stride,
                                 const double2 *__restrict__ tw, int logn, int chunk_log,
                                 volatile uint32_t *flag, uint32_t token) {
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
        k2_smem[k2_pos(k2_bitrev(q, chunk_log), chunk_log)] =
            src < n_in ? in[src] : make_double2(0.0, 0.0);   // zero pad, src/fft.rs:36
    }
    __syncthreads();
    auto ld = [&](size_t e) { return k2_smem[k2_pos((uint32_t)e, chunk_log)]; };
    auto st = [&](size_t e, double2 v) { k2_smem[k2_pos((uint32_t)e, chunk_log)] = v; };
    int lv = 1;
    const int r0 = chunk_log % 3;
    if (r0 == 1) {
        for (uint32_t it = threadIdx.x; it < cn / 2; it += blockDim.x) k2_levels<1>(tw, logn, lv, it, ld, st);
        lv += 1;
        __syncthreads();
    } else if (r0 == 2) {
        for (uint32_t it = threadIdx.x; it < cn / 4; it += blockDim.x) k2_levels<2>(tw, logn, lv, it, ld, st);
        lv += 2;
        __syncthreads();
    }
    for (; lv <= chunk_log; lv += 3) {
        for (uint32_t it = threadIdx.x; it < cn / 8; it += blockDim.x) k2_levels<3>(tw, logn, lv, it, ld, st);
        __syncthreads();
    }
    for (uint32_t p = threadIdx.x; p < cn; p += blockDim.x) out[base + p] = k2_smem[k2_pos(p, chunk_log)];
    if (flag) {
        __threadfence_system();          // this thread's stores are visible to the host before ...
        __syncthreads();
        if (threadIdx.x == 0) *flag = token;   // ... the token is
    }
}

// Levels lv0 .. lv0+R-1 above the chunk size in global memory, in place.  One work item (2^R elements) per thread.
template <int R>
__global__ void k2_c2c_f64_levels(double2 *__restrict__ data, const double2 *__restrict__ tw, int logn, int lv0) {
    const size_t item = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= ((size_t)1 << (logn - R))) return;
    k2_levels<R>(tw, logn, lv0, item, [&](size_t e) { return data[e]; }, [&](size_t e, double2 v) { data[e] = v; });
}

}  // namespace kopek
