// K3 -- large single complex FFT (n = N1 x N2 up to 2^24) as a four-step decomposition with
// TMA-staged tiles (sm_100a).  BASELINE.json configs[3]: "single large 2^24-point FFT".
//
// The reference computes any length with one recursive call (src/fft.rs:6-41); for a 2^24-point
// transform its two 256 MiB ping-pong buffers are walked with strides up to 2^23 elements.  Here:
//
//   n = n1 N2 + n2,  k = k1 + N1 k2
//   pass 1 (one CTA per tile of 2 columns n2):  A[k1][n2] = W_n^{n2 k1} * sum_n1 x[n1 N2 + n2] W_N1^{n1 k1}
//   pass 2 (one CTA per tile of 4 rows k1):     X[k1 + N1 k2] = sum_n2 A[k1][n2] W_N2^{n2 k2}
//
// Two HBM passes (the minimum for a transform that does not fit on chip).  Every global access is a
// TMA tile: pass 1 loads a [N1 x 2] column tile with 2-D boxes (16-byte rows, two CTAs per SM) and stores its
// result TILE-MAJOR -- mid[k1/4][n2/2][k1%4][n2%2], 64-byte segments -- so that pass 2 loads ONE contiguous
// block per CTA (1-D bulk copies) and stores [N2 x 4] tiles of the natural-order output.  Both transposes and the
// inter-step twiddle are fused into the passes; there is no standalone transpose kernel.
// The sub-transforms are Stockham radix-4/8/16 in shared memory with the four tile columns interleaved
// (column index fastest across lanes), which keeps every shared access contiguous.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "common.cuh"
// complex add / subtract as ONE packed instruction (add/sub.rn.f32x2 -> SASS FADD2) in this translation unit: the two
// passes run under a 64-register cap with a quarter of their instructions in address arithmetic, and 11 % fewer
// issue slots are worth 3.4 % at 2^24 (0.1464 -> 0.1415 ms).  The batched kernels gain nothing from it (measured,
// profiles/r1_f32x2_sweep.txt), so they keep the scalar form.
#define KOPEK_F32X2
#include "fft_batched.cuh"
#include "lane_common.cuh"

namespace kopek {

// Resident CTAs per SM asked of the compiler for the small tiles (N <= 1024: n <= 2^21).  There a pass is one or two
// rounds of tiles (2^20: 512 pass-1 tiles on 444 slots, 256 pass-2 tiles on 148), and asking for four / two CTAs per SM
// so that every tile finds a slot in the first round was measured (profiles/r2_k3_small_sizes.txt): 128 registers
// instead of 143 / 188 cost 24-40 bytes of spills and the passes get SLOWER (2^20 in a CUDA graph 14.0 -> 14.9 us,
// 2^19 9.4 -> 11.6 us) -- the tile's own latency chain, not the second round, is what a small transform waits for.
#ifndef KOPEK_K3_MINB1
#define KOPEK_K3_MINB1 2
#endif
#ifndef KOPEK_K3_MINB2
#define KOPEK_K3_MINB2 1
#endif
template <int M> struct K3Cfg;
template <> struct K3Cfg<256>  { static constexpr int R0 = 16, R1 = 16, R2 = 1; };
template <> struct K3Cfg<512>  { static constexpr int R0 = 8,  R1 = 8,  R2 = 8; };
template <> struct K3Cfg<1024> { static constexpr int R0 = 4,  R1 = 16, R2 = 16; };
template <> struct K3Cfg<2048> { static constexpr int R0 = 8,  R1 = 16, R2 = 16; };
template <> struct K3Cfg<4096> { static constexpr int R0 = 16, R1 = 16, R2 = 16; };

__host__ __device__ constexpr int k3_swz(int q) { return q ^ ((q >> 4) & 15); }

// One Stockham stage on 4 interleaved columns: element (q, c) lives at buf[pos(q) * 4 + c].
// READ_LINEAR / WRITE_LINEAR select the dense layout TMA produced / consumes; otherwise the swizzled one.
// FIRST_MAP remaps the logical index of the very first read (pass 2: tile-major block order).
// XTW (last stage of pass 1 only): the inter-step twiddle W_n^{n2 q} is applied to output q in registers just
// before the final write (two-level table: W^m = W^{4096 (m >> 12)} W^{m & 4095}), saving a pass over the tile.
struct K3NoOut {
    static constexpr bool enabled = false;
    __device__ __forceinline__ void reads_done() const {}
    __device__ __forceinline__ void store(int, bool, float2) const {}
};

// OUT (last stage only): instead of writing the result back to shared memory, `out.reads_done()` runs right after the
// barrier that ends the stage's reads (the tile buffer is free from here on: the caller starts the TMA load of the
// NEXT tile into it) and every output q goes straight to global memory through `out.store(q, v)`.
template <int COLS, int RAD, int NS, int M, int T, bool READ_LINEAR, bool WRITE_LINEAR, bool TILE_MAJOR_IN,
          bool XTW = false, class OUT = K3NoOut>
__device__ __forceinline__ void k3_stage(float2 *buf, const float2 *tw, int tid, int c, uint32_t n2 = 0,
                                         const float2 *tw_hi = nullptr, const float2 *tw_lo = nullptr,
                                         const OUT &out = OUT()) {
    constexpr int NB = 16 / RAD;
    float2 d[16];
    // The four table lookups of the inter-step twiddle do not depend on the tile.  For M < 4096 they are issued before
    // the shared-memory reads, so that their L2 latency (the tables do not fit the L1 the tile buffers leave) hides
    // behind the butterflies -- which also frees the compiler from the spills it had at 128 registers (M = 2048: STACK
    // 40 -> 0; 2^22 41.4 -> 40.5 us, 2^23 79.9 -> 77.8 us).  At M = 4096 (64 registers) the four extra live values
    // cost more than the latency they hide (2^24 124.3 -> 126.5 us in a graph): looked up late there.
    constexpr bool XTW_EARLY = XTW && M < 4096;
    float2 x_b0[NB], x_st[NB];
    if constexpr (XTW_EARLY) {
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int j = tid + T * i, kk = j & (NS - 1), base = (j - kk) * RAD + kk;
            const uint32_t m0 = n2 * (uint32_t)base, ms = n2 * (uint32_t)NS;
            x_b0[i] = cmul(__ldg(tw_hi + (m0 >> 12)), __ldg(tw_lo + (m0 & 4095u)));
            x_st[i] = cmul(__ldg(tw_hi + (ms >> 12)), __ldg(tw_lo + (ms & 4095u)));
        }
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        int j = tid + T * i;
        int kk = j & (NS - 1);
#pragma unroll
        for (int t = 0; t < RAD; ++t) {
            int q = j + t * (M / RAD);
            int pos;
            if constexpr (TILE_MAJOR_IN) pos = (q >> 1) * (2 * COLS) + (c << 1) + (q & 1);   // mid tile: [q/2][c][q%2]
            else pos = (READ_LINEAR ? q : k3_swz(q)) * COLS + c;
            float2 v = buf[pos];
            if (NS > 1 && t > 0) v = cmul(v, tw[(t - 1) * NS + kk]);           // stage table (shared), t-major
            d[i * RAD + t] = v;
        }
    }
    if constexpr (OUT::enabled) {
        __syncthreads();
        out.reads_done();
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) Dft<RAD>::run(d + i * RAD);
    if constexpr (!OUT::enabled) __syncthreads();
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        int j = tid + T * i;
        int kk = j & (NS - 1);
        int base = (j - kk) * RAD + kk;
        if constexpr (XTW) {
            // outputs q = base + k NS: W_n^{n2 q} = W_n^{n2 base} (W_n^{n2 NS})^k -- two table lookups per
            // butterfly, the powers by squaring/multiplying (depth <= 4, error ~4 ulp)
            float2 b0, st;
            if constexpr (XTW_EARLY) {
                b0 = x_b0[i], st = x_st[i];
            } else {
                const uint32_t m0 = n2 * (uint32_t)base, ms = n2 * (uint32_t)NS;
                b0 = cmul(__ldg(tw_hi + (m0 >> 12)), __ldg(tw_lo + (m0 & 4095u)));
                st = cmul(__ldg(tw_hi + (ms >> 12)), __ldg(tw_lo + (ms & 4095u)));
            }
            float2 pw[RAD];
            pw[0] = make_float2(1.0f, 0.0f);
            pw[1] = st;
#pragma unroll
            for (int k = 2; k < RAD; ++k) pw[k] = (k & 1) ? cmul(pw[k - 1], st) : cmul(pw[k / 2], pw[k / 2]);
            d[i * RAD] = cmul(d[i * RAD], b0);
#pragma unroll
            for (int k = 1; k < RAD; ++k) d[i * RAD + k] = cmul(d[i * RAD + k], cmul(b0, pw[k]));
        }
#pragma unroll
        for (int k = 0; k < RAD; ++k) {
            int q = base + k * NS;
            if constexpr (OUT::enabled) out.store(q, 2 * k >= RAD, d[i * RAD + k]);
            else buf[(WRITE_LINEAR ? q : k3_swz(q)) * COLS + c] = d[i * RAD + k];
        }
    }
    if constexpr (!OUT::enabled) __syncthreads();
}

// M-point FFT of the COLS interleaved columns in `buf`; the last stage hands its outputs to `out` (global memory).
template <int COLS, int M, bool TILE_MAJOR_IN, bool XTW, class OUT>
__device__ __forceinline__ void k3_fft_tile(float2 *buf, const float2 *tw, int tid, int c, const OUT &out,
                                            uint32_t n2 = 0, const float2 *tw_hi = nullptr,
                                            const float2 *tw_lo = nullptr) {
    using C = K3Cfg<M>;
    constexpr int T = M / 16;
    if constexpr (C::R2 > 1) {
        k3_stage<COLS, C::R0, 1, M, T, true, false, TILE_MAJOR_IN>(buf, tw, tid, c);
        k3_stage<COLS, C::R1, C::R0, M, T, false, false, false>(buf, tw, tid, c);
        k3_stage<COLS, C::R2, C::R0 * C::R1, M, T, false, true, false, XTW, OUT>(buf, tw + (C::R1 - 1) * C::R0, tid, c,
                                                                                  n2, tw_hi, tw_lo, out);
    } else {
        k3_stage<COLS, C::R0, 1, M, T, true, false, TILE_MAJOR_IN>(buf, tw, tid, c);
        k3_stage<COLS, C::R1, C::R0, M, T, false, true, false, XTW, OUT>(buf, tw, tid, c, n2, tw_hi, tw_lo, out);
    }
}

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int x, int y, const void *src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(map), "r"(x), "r"(y), "r"(smem_u32(src)) : "memory");
}
__device__ __forceinline__ void tma_store_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---- L2 residency control (round 2) -------------------------------------------------------------------------------
// `mid` is written by pass 1 and read once by pass 2.  The part of it that still sits in L2 when pass 2 starts needs
// neither the HBM read nor -- if pass 2 drops the lines after reading them -- the HBM write-back.  Pass 1 therefore
// stores the rows of mid it wants kept (all of them while mid fits, else the upper half of k1) with an evict_last
// policy and the rest evict_first; pass 2 walks the tiles from the top, so that the kept rows come first, and
// discards their lines once the tile is in shared memory (discard.global.L2: the line is dropped without write-back;
// `mid` is dead after this read).  The INPUT stream keeps the normal policy: adjacent two-column tiles share 32-byte
// sectors, and an evict_first line was gone before the neighbouring CTA read its half (+20 % DRAM reads, measured).
// Measured (profiles/r2_k3_keep_sweep.txt): 2^22 0.0528 -> 0.0487 ms, 2^23 0.0841 -> 0.0785 ms, 2^24 0.1357 -> 0.1333 ms.
// -DKOPEK_K3_NO_HINTS (tuning builds) compiles all of it away: plain stores, plain TMA loads, no discard.
template <int KIND> __device__ __forceinline__ uint64_t l2_policy() {       // 1 evict_first, 2 evict_last
    uint64_t p = 0;
#ifndef KOPEK_K3_NO_HINTS
    // not volatile: a policy is a constant the compiler may rematerialise instead of keeping it in registers
    if constexpr (KIND == 1) asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    else asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
#endif
    return p;
}
__device__ __forceinline__ void st_policy(float2 *ptr, float2 v, uint64_t pol) {
#ifndef KOPEK_K3_NO_HINTS
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(ptr), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
#else
    *ptr = v;
#endif
}
__device__ __forceinline__ void l2_discard_line(const void *ptr) {
#ifndef KOPEK_K3_NO_HINTS
    asm volatile("discard.global.L2 [%0], 128;" ::"l"(ptr) : "memory");
#endif
}
__device__ __forceinline__ void tma_load_1d_policy(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint64_t pol) {
#ifndef KOPEK_K3_NO_HINTS
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
#else
    tma_load_1d(dst_smem, src_gmem, bytes, bar);
#endif
}

// Programmatic dependent launch: pass 2 is launched while pass 1 still runs; its CTAs become resident as pass-1 CTAs
// retire, load their stage tables into shared memory and initialise their barrier, and only then wait for pass 1 to
// have completed (griddepcontrol.wait) before the first read of `mid`.  Hides pass 2's prologue and launch latency
// behind pass 1's tail: what a 2^20-point transform (two ~9 us kernels) spends a fifth of its time on.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_primary() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

struct K3Tables {
    const float2 *tw_sub1;   // per-stage t-major twiddles of the N1-point sub-transform (k3_stage_twiddles)
    const float2 *tw_sub2;   // same for N2
    const float2 *tw_hi;     // W_n^{4096 h}, h < n/4096 (two-level inter-step twiddle)
    const float2 *tw_lo;     // W_n^{l},      l < 4096
};

template <int N> __device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }

// Both passes are PERSISTENT CTAs looping over their tiles.  A tile is loaded by TMA (pass 1: 2-D boxes of the
// strided column tile; pass 2: 1-D bulk copies of one contiguous block), transformed in place in shared memory, and
// the LAST stage stores its outputs from registers straight to global memory.  The tile buffer is therefore free as
// soon as the last stage has read it: the load of the CTA's next tile is issued at that point and overlaps the last
// butterflies, the inter-step twiddle and the stores (the first version wrote the result back to shared memory,
// stored it with TMA and could only reload a box once that store had drained: 0.162 ms at 2^24).
// tools/micro/tma_rows.cu measured the memory side alone on this matrix: TMA loads of 16 / 32 / 64-byte rows run at
// 4.0 / 4.9 / 5.8 TB/s, TMA stores at 2.7 / 4.6 / 4.9 TB/s, plain 64-byte-run stores at 5.3 TB/s.

// pass 1 output: thread (tid, c) holds A[k1][n2 = 2 s + c]; mid is tile-major [k1/4][n2/2][k1%4][n2%2], so the 16
// lanes x 2 columns of a warp write four 64-byte segments per store instruction
// KEEP: which rows k1 of mid pass 1 asks L2 to hold until pass 2 has read them -- 0 none, 1 the upper half (the flag
// `upper` is a compile-time property of the butterfly output after unrolling: no select, no compare), 2 all of them
template <int KEEP> struct K3Out1 {
    static constexpr bool enabled = true;
    float2 *dst;               // mid + s * 8 + c
    size_t seg_stride;         // (N2 / 2) * 8 complex between k1/4 blocks
    uint64_t *bar;
    const CUtensorMap *in_map;
    float2 *buf;
    int next_tile, n1;
    uint64_t pol_keep, pol_stream;
    __device__ __forceinline__ void reads_done() const {
        if (threadIdx.x == 0 && next_tile >= 0) {
            mbar_expect_tx(bar, n1 * 16);
            for (int r0 = 0; r0 < n1; r0 += 256) tma_load_2d(buf + r0 * 2, in_map, 4 * next_tile, r0, bar);
        }
    }
    __device__ __forceinline__ void store(int k1, bool upper, float2 v) const {
        st_policy(dst + (size_t)(k1 >> 2) * seg_stride + (k1 & 3) * 2, v,
                  (KEEP == 2 || (KEEP == 1 && upper)) ? pol_keep : pol_stream);
    }
};

// ---- pass 1: column tiles s (columns 2s, 2s+1), all n1.  Two-column tiles (N1 x 16 B = 64 KB at N1 = 4096)
// let TWO CTAs share an SM, so one transforms while the other waits on shared memory or on its TMA traffic.
template <int N1, int KEEP>
__global__ void __launch_bounds__(N1 / 8, N1 <= 1024 ? KOPEK_K3_MINB1 : 2)
k3_pass1(const __grid_constant__ CUtensorMap in_map, float2 *__restrict__ mid, K3Tables tb, int n_tiles,
         const float2 *__restrict__ in) {
    extern __shared__ __align__(128) unsigned char k3_smem[];
    float2 *buf = reinterpret_cast<float2 *>(k3_smem);                 // N1 x 2 complex, dense
    float2 *s_tw = reinterpret_cast<float2 *>(k3_smem + (size_t)N1 * 16);   // stage twiddles, N1 entries
    uint64_t *bar = reinterpret_cast<uint64_t *>(k3_smem + (size_t)N1 * 24);
    pdl_launch_dependents();                                  // pass 2 may start its prologue as SMs free up
    for (int i = threadIdx.x; i < N1; i += N1 / 8) s_tw[i] = tb.tw_sub1[i];
    const uint64_t pol_keep = l2_policy<2>(), pol_stream = l2_policy<1>();
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // pass 1 is itself launched with programmatic serialization: everything above (tables are never written by a
    // kernel) may overlap the tail of whatever precedes it in the stream -- typically pass 2 of the previous transform;
    // the input and the workspace are touched only from here on
    pdl_wait_primary();
    if (threadIdx.x == 0) {
        if ((int)blockIdx.x < n_tiles) {
            mbar_expect_tx(bar, N1 * 16);
            for (int r0 = 0; r0 < N1; r0 += 256) tma_load_2d(buf + r0 * 2, &in_map, 4 * blockIdx.x, r0, bar);
        }
    }
    __syncthreads();
    const int c = threadIdx.x & 1, tid = threadIdx.x >> 1;
    uint32_t phase = 0;
    for (int s = blockIdx.x; s < n_tiles; s += gridDim.x) {
        mbar_wait(bar, phase);
        phase ^= 1;
        const int nxt = s + gridDim.x;
        // The next tile's TMA load can only be issued when this tile's last stage has read the buffer, and a quarter of
        // the pass is spent waiting for it (ncu source view: the mbarrier wait).  With `in` set its DRAM latency is paid
        // now instead: the eight tiles that share the 128-byte lines of a row are in flight on eight CTAs at the same
        // time, so each of them asks L2 for the lines of one row in eight (one line per thread).  The host enables
        // it where it was measured to pay (profiles/r2_k3_small_sizes.txt): 2^23, -6 %; at 2^24 the prefetched
        // lines compete with the part of `mid` that is kept in L2, at <= 2^22 the benchmark's input is L2-resident.
        if (in != nullptr && nxt < n_tiles)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(in + ((size_t)((nxt & 7) + 8 * (int)threadIdx.x) * (2 * (size_t)n_tiles) +
                                                               (size_t)(nxt >> 3) * 16)));
        K3Out1<KEEP> out{mid + (size_t)s * 8 + c, (size_t)n_tiles * 8, bar, &in_map, buf, nxt < n_tiles ? nxt : -1, N1,
                         pol_keep, pol_stream};
        // sub-transform + inter-step twiddle W_n^{n2 k1}, n2 = 2s + c (applied in registers before the stores)
        k3_fft_tile<2, N1, false, true>(buf, s_tw, tid, c, out, 2u * s + c, tb.tw_hi, tb.tw_lo);
    }
}

// pass 2 output: thread (tid, c) holds X[k1 + N1 k2], k1 = 4 u + c: the 8 lanes x 4 rows of a warp write eight
// 32-byte runs per store instruction
struct K3Out2 {
    static constexpr bool enabled = true;
    float2 *dst;               // out + 4 u + c
    size_t n1;
    uint64_t *bar;
    const char *next_src;      // next tile's contiguous block of mid, or nullptr
    unsigned char *smem;
    int bytes, chunk;
    uint64_t pol_load, pol_store;
    __device__ __forceinline__ void reads_done() const {
        if (threadIdx.x == 0 && next_src) {
            mbar_expect_tx(bar, bytes);
            for (int off = 0; off < bytes; off += chunk) tma_load_1d_policy(smem + off, next_src + off, chunk, bar, pol_load);
        }
    }
    __device__ __forceinline__ void store(int k2, bool, float2 v) const { st_policy(dst + (size_t)k2 * n1, v, pol_store); }
};

// ---- pass 2: row tiles u (rows k1 = 4u..4u+3), all n2; block mid[u] is contiguous
template <int N2>
__global__ void __launch_bounds__(N2 / 4, N2 <= 1024 ? KOPEK_K3_MINB2 : 1)
k3_pass2(const float2 *__restrict__ mid, float2 *__restrict__ outp, K3Tables tb, int n_tiles, int keep_from_tile) {
    extern __shared__ __align__(128) unsigned char k3_smem[];
    float2 *buf = reinterpret_cast<float2 *>(k3_smem);
    float2 *s_tw = reinterpret_cast<float2 *>(k3_smem + (size_t)N2 * 32);
    uint64_t *bar = reinterpret_cast<uint64_t *>(k3_smem + (size_t)N2 * 40);
    for (int i = threadIdx.x; i < N2; i += N2 / 4) s_tw[i] = tb.tw_sub2[i];
    constexpr int CH = N2 * 32 < 16384 ? N2 * 32 : 16384;              // 1-D bulk copy granule
    const uint64_t pol_stream = l2_policy<1>(), pol_normal = l2_policy<2>();    // kept tiles: stay until read, then dropped
    // tiles are walked from the top: u = n_tiles - 1 - v, so the rows pass 1 asked L2 to keep are read first
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    pdl_wait_primary();                                       // mid is complete and visible from here on
    if (threadIdx.x == 0) {
        if ((int)blockIdx.x < n_tiles) {
            const int u0 = n_tiles - 1 - (int)blockIdx.x;
            mbar_expect_tx(bar, N2 * 32);
            const char *src = reinterpret_cast<const char *>(mid) + (size_t)u0 * N2 * 32;
            for (int off = 0; off < N2 * 32; off += CH)
                tma_load_1d_policy(k3_smem + off, src + off, CH, bar, u0 >= keep_from_tile ? pol_normal : pol_stream);
        }
    }
    __syncthreads();
    const int c = threadIdx.x & 3, tid = threadIdx.x >> 2;
    const size_t n1 = (size_t)n_tiles * 4;
    uint32_t phase = 0;
    for (int v = blockIdx.x; v < n_tiles; v += gridDim.x) {
        const int u = n_tiles - 1 - v;
        if (v + (int)gridDim.x >= n_tiles) pdl_launch_dependents();   // last tile: the next kernel may bring its prologue in
        mbar_wait(bar, phase);
        phase ^= 1;
        // the tile is in shared memory: a kept tile's lines are dead in L2 (one line per thread: N2 * 32 / 128 = N2 / 4)
        if (u >= keep_from_tile)
            l2_discard_line(reinterpret_cast<const char *>(mid) + (size_t)u * N2 * 32 + (size_t)threadIdx.x * 128);
        const int nv = v + gridDim.x, nu = n_tiles - 1 - nv;
        K3Out2 out{outp + 4 * (size_t)u + c, n1, bar,
                   nv < n_tiles ? reinterpret_cast<const char *>(mid) + (size_t)nu * N2 * 32 : nullptr, k3_smem,
                   N2 * 32, CH, nu >= keep_from_tile ? pol_normal : pol_stream, pol_stream};
        k3_fft_tile<4, N2, true, false>(buf, s_tw, tid, c, out);
    }
}

// ---- pass 2, N2 = 4096, as TWO stages of radix 64 (64 x 64) instead of three of radix 16.
// A thread carries 64 values (128 registers) through a register-resident 64-point transform, so the tile crosses
// shared memory once between the stages instead of twice, every shared access is base + compile-time offset (the
// exchange is PADDED -- element q sits at q + (q >> 6) -- where the radix-16 path needs an XOR per access), and the
// stage table is read 63 times per 64 points.  256 threads per CTA (64 per row, four rows interleaved).
constexpr int K3R_TILE_F2 = 4096 * 4;                    // dense tile as it lands
constexpr int K3R_EX_F2 = (4096 + 64) * 4;               // padded exchange, in place over the landing buffer
constexpr int K3R_TW_F2 = 63 * 64;                       // [t-1][kk] = W4096^{kk t}
constexpr int K3R_SMEM = K3R_EX_F2 * 8 + K3R_TW_F2 * 8 + 16;

__global__ void __launch_bounds__(256, 1)
k3_pass2_r64(const float2 *__restrict__ mid, float2 *__restrict__ outp, const float2 *__restrict__ tw_r64, int n_tiles,
             int keep_from_tile) {
    extern __shared__ __align__(128) unsigned char k3_smem[];
    float2 *buf = reinterpret_cast<float2 *>(k3_smem);
    float2 *s_tw = buf + K3R_EX_F2;
    uint64_t *bar = reinterpret_cast<uint64_t *>(s_tw + K3R_TW_F2);
    for (int i = threadIdx.x; i < K3R_TW_F2; i += 256) s_tw[i] = tw_r64[i];
    constexpr int BYTES = K3R_TILE_F2 * 8, CH = 16384;
    const uint64_t pol_stream = l2_policy<1>(), pol_normal = l2_policy<2>();
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    pdl_wait_primary();                                       // mid is complete and visible from here on
    if (threadIdx.x == 0) {
        if ((int)blockIdx.x < n_tiles) {
            const int u0 = n_tiles - 1 - (int)blockIdx.x;              // tiles from the top: the kept rows first
            mbar_expect_tx(bar, BYTES);
            const char *src = reinterpret_cast<const char *>(mid) + (size_t)u0 * BYTES;
            for (int off = 0; off < BYTES; off += CH)
                tma_load_1d_policy(k3_smem + off, src + off, CH, bar, u0 >= keep_from_tile ? pol_normal : pol_stream);
        }
    }
    __syncthreads();
    const int c = threadIdx.x & 3, tid = threadIdx.x >> 2;
    const size_t n1 = (size_t)n_tiles * 4;
    const float2 *rdA = buf + (tid >> 1) * 8 + 2 * c + (tid & 1);    // tile-major [n2/2][k1%4][n2%2], n2 = tid + 64 t
    float2 *wrA = buf + (65 * tid) * 4 + c;                           // q = 64 tid + k  ->  65 tid + k
    const float2 *rdB = buf + tid * 4 + c;                            // q = tid + 64 t  ->  tid + 65 t
    const float2 *twB = s_tw + tid;
    uint32_t phase = 0;
    for (int tv = blockIdx.x; tv < n_tiles; tv += gridDim.x) {
        const int u = n_tiles - 1 - tv;
        if (tv + (int)gridDim.x >= n_tiles) pdl_launch_dependents();  // last tile: the next kernel may bring its prologue in
        mbar_wait(bar, phase);
        phase ^= 1;
        if (u >= keep_from_tile) {                                    // the tile has landed: its 1024 lines are dead in L2
            const char *line = reinterpret_cast<const char *>(mid) + (size_t)u * BYTES + (size_t)threadIdx.x * 128;
#pragma unroll
            for (int i = 0; i < BYTES / 128 / 256; ++i) l2_discard_line(line + (size_t)i * 256 * 128);
        }
        float2 v[64];
#pragma unroll
        for (int t = 0; t < 64; ++t) v[t] = rdA[256 * t];
        __syncthreads();                                              // the exchange overwrites the landing buffer
        Dft<64>::run(v);
#pragma unroll
        for (int k = 0; k < 64; ++k) wrA[4 * k] = v[k];
        __syncthreads();
#pragma unroll
        for (int t = 0; t < 64; ++t) {
            float2 x = rdB[260 * t];
            v[t] = t ? cmul(x, twB[64 * (t - 1)]) : x;
        }
        __syncthreads();
        // the buffer is free: the next tile lands while this one is finished and stored from registers
        const int nxt = tv + gridDim.x;
        if (threadIdx.x == 0 && nxt < n_tiles) {
            const int nu = n_tiles - 1 - nxt;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar, BYTES);
            const char *src = reinterpret_cast<const char *>(mid) + (size_t)nu * BYTES;
            for (int off = 0; off < BYTES; off += CH)
                tma_load_1d_policy(k3_smem + off, src + off, CH, bar, nu >= keep_from_tile ? pol_normal : pol_stream);
        }
        Dft<64>::run(v);
        // X[k1 + N1 k2], k1 = 4 u + c, k2 = tid + 64 k: the 8 lanes x 4 rows of a warp write eight 32-byte runs
        float2 *dst = outp + 4 * (size_t)u + c + (size_t)tid * n1;
        const size_t step = 64 * n1;
#pragma unroll
        for (int k = 0; k < 64; ++k) { st_policy(dst, v[k], pol_stream); dst += step; }
    }
}

// ------------------------------------------------------------------------- host ----
typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                              const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_fn get_encode() {
    static encode_fn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<encode_fn>(p);
    });
    return fn;
}

// 2-D f32 tensor map: `rows` rows of `row_floats` floats, box = box_floats x 256 rows.
static bool make_map(CUtensorMap *m, const void *base, uint64_t row_floats, uint64_t rows, uint32_t box_floats,
                     uint32_t box_rows) {
    encode_fn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {row_floats, rows};
    cuuint64_t strides[1] = {row_floats * sizeof(float)};
    cuuint32_t box[2] = {box_floats, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct K3Plan {
    uint64_t n = 0;
    int device = -1;
    int n1 = 0, n2 = 0;
    float2 *tw1 = nullptr, *tw2 = nullptr, *hi = nullptr, *lo = nullptr;
    float2 *tw2r = nullptr;    // N2 = 4096: stage table of the two-stage radix-64 pass 2
    // launch geometry, fixed when the plan is made (function attributes and occupancy are per device, like the plan):
    // a call then costs one tensor-map encode, the workspace allocation and the two launches on the host
    int sms = 0, grid1[3] = {0, 0, 0}, grid2 = 0;
};
// Tables only: one entry per (n, device), at most 9 lengths per device, kept for the life of the process, so a plan
// handed out under the lock stays valid however many threads and streams call in.  The pass-1 -> pass-2 workspace
// `mid` is NOT part of the plan: it is allocated per call, stream-ordered on the caller's stream, from a per-device
// pool that keeps its memory between calls (k3_workspace_pool) -- two calls on different streams never share it.
static std::mutex g_k3_mu;
static std::vector<K3Plan> g_k3_plans;
static cudaMemPool_t g_k3_pool[64];

static kopek_status k3_workspace_pool(int device, cudaMemPool_t *out) {
    if (device < 0 || device >= 64) return fail(KOPEK_ERR_INVALID, "device %d out of range", device);
    std::lock_guard<std::mutex> lk(g_k3_mu);
    if (!g_k3_pool[device]) {
        cudaMemPoolProps props{};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaMemPool_t pool = nullptr;
        KOPEK_CUDA(cudaMemPoolCreate(&pool, &props));
        uint64_t keep = ~0ull;                    // never trim at synchronisation points: the next call reuses the block
        KOPEK_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        g_k3_pool[device] = pool;
    }
    *out = g_k3_pool[device];
    return KOPEK_OK;
}

// Stage tables of the M-point Stockham sub-transform: block s holds tw[(t-1) NS + kk] = W_{NS RAD}^{kk t}.
template <int M> static void stage_twiddles_t(std::vector<float2> &v) {
    using C = K3Cfg<M>;
    v.assign(M, make_float2(1.0f, 0.0f));
    const int rad[2] = {C::R1, C::R2}, ns[2] = {C::R0, C::R0 * C::R1};
    int off = 0;
    for (int s = 0; s < 2; ++s) {
        if (rad[s] <= 1) break;
        for (int t = 1; t < rad[s]; ++t)
            for (int kk = 0; kk < ns[s]; ++kk) {
                double th = -2.0 * M_PI * (double)(kk * t) / (double)(ns[s] * rad[s]);
                v[off + (t - 1) * ns[s] + kk] = make_float2((float)cos(th), (float)sin(th));
            }
        off += (rad[s] - 1) * ns[s];
    }
}
static void k3_stage_twiddles(int m, std::vector<float2> &v) {
    switch (m) {
        case 256: stage_twiddles_t<256>(v); break;
        case 512: stage_twiddles_t<512>(v); break;
        case 1024: stage_twiddles_t<1024>(v); break;
        case 2048: stage_twiddles_t<2048>(v); break;
        default: stage_twiddles_t<4096>(v); break;
    }
}

static void fill_twiddles(std::vector<float2> &v, uint64_t count, uint64_t step, uint64_t n) {
    v.resize(count);
    for (uint64_t i = 0; i < count; ++i) {
        double th = -2.0 * M_PI * (double)(i * step) / (double)n;
        v[i] = make_float2((float)cos(th), (float)sin(th));
    }
}

static cudaError_t k3_prepare(K3Plan &p);

static kopek_status k3_get_plan(uint64_t n, int device, K3Plan *out) {
    std::lock_guard<std::mutex> lk(g_k3_mu);
    for (const K3Plan &p : g_k3_plans)
        if (p.n == n && p.device == device) { *out = p; return KOPEK_OK; }
    int lg = 0;
    while (((uint64_t)1 << lg) < n) ++lg;
    K3Plan p;
    p.n = n; p.device = device;
    p.n1 = 1 << (lg / 2);
    p.n2 = (int)(n / p.n1);
    std::vector<float2> t1, t2, hi, lo;
    k3_stage_twiddles(p.n1, t1);
    k3_stage_twiddles(p.n2, t2);
    fill_twiddles(hi, n / 4096, 4096, n);
    fill_twiddles(lo, 4096, 1, n);
    auto up = [](float2 **d, const std::vector<float2> &h) {
        return cudaMalloc(d, h.size() * sizeof(float2)) == cudaSuccess &&
               cudaMemcpy(*d, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice) == cudaSuccess;
    };
    std::vector<float2> t2r;
    if (p.n2 == 4096) {
        t2r.resize(K3R_TW_F2);
        for (int t = 1; t < 64; ++t)
            for (int kk = 0; kk < 64; ++kk) {
                double th = -2.0 * M_PI * (double)(kk * t) / 4096.0;
                t2r[(t - 1) * 64 + kk] = make_float2((float)cos(th), (float)sin(th));
            }
    }
    if (!up(&p.tw1, t1) || !up(&p.tw2, t2) || !up(&p.hi, hi) || !up(&p.lo, lo) || (!t2r.empty() && !up(&p.tw2r, t2r))) {
        cudaFree(p.tw1); cudaFree(p.tw2); cudaFree(p.hi); cudaFree(p.lo); cudaFree(p.tw2r);
        return fail(KOPEK_ERR_NOMEM, "four-step tables for n=%llu: %s", (unsigned long long)n,
                    cudaGetErrorString(cudaGetLastError()));
    }
    cudaError_t pe = k3_prepare(p);
    if (pe != cudaSuccess) {
        cudaFree(p.tw1); cudaFree(p.tw2); cudaFree(p.hi); cudaFree(p.lo); cudaFree(p.tw2r);
        cudaGetLastError();
        return fail(KOPEK_ERR_CUDA, "four-step kernels for n=%llu: %s", (unsigned long long)n, cudaGetErrorString(pe));
    }
    g_k3_plans.push_back(p);
    *out = p;
    return KOPEK_OK;
}

// launch with programmatic stream serialization: the kernel may begin before the previous one in the stream has
// finished and must call griddepcontrol.wait before it touches that kernel's results (pdl_wait_primary above)
template <class... P, class... A>
static cudaError_t launch_pdl(void (*kern)(P...), int grid, int block, size_t smem, cudaStream_t st, A... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<P>(args)...);
}

// persistent grid: as many CTAs as can be resident (occupancy x SMs), at most one per tile
template <class K> static int k3_grid(K kern, int threads, int smem, int tiles, int sms) {
    int nb = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, smem) != cudaSuccess || nb < 1) nb = 1;
    long long g = (long long)nb * sms;
    return (int)(g < tiles ? g : tiles);
}
// prepare_*: opt the kernel into its shared-memory size on the current device and size the persistent grid (once per plan)
template <int N1, int KEEP> static cudaError_t prepare_pass1k(int tiles, int sms, int *grid) {
    const int smem = N1 * 24 + 16;
    cudaError_t e = cudaFuncSetAttribute(k3_pass1<N1, KEEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) *grid = k3_grid(k3_pass1<N1, KEEP>, N1 / 8, smem, tiles, sms);
    return e;
}
template <int N1> static cudaError_t prepare_pass1(int tiles, int sms, int (&grid)[3]) {
    cudaError_t e = prepare_pass1k<N1, 0>(tiles, sms, &grid[0]);
    if (e == cudaSuccess) e = prepare_pass1k<N1, 1>(tiles, sms, &grid[1]);
    if (e == cudaSuccess) e = prepare_pass1k<N1, 2>(tiles, sms, &grid[2]);
    return e;
}
template <int N2> static cudaError_t prepare_pass2(int tiles, int sms, int *grid) {
    const int smem = N2 * 40 + 16;
    cudaError_t e = cudaFuncSetAttribute(k3_pass2<N2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) *grid = k3_grid(k3_pass2<N2>, N2 / 4, smem, tiles, sms);
    return e;
}
static const bool k3_chain = !(getenv("KOPEK_K3_CHAIN") && atoi(getenv("KOPEK_K3_CHAIN")) == 0);   // tuning switch
static const bool k3_use_r64 = !(getenv("KOPEK_K3_R64") && atoi(getenv("KOPEK_K3_R64")) == 0);   // tuning switch
static cudaError_t k3_prepare(K3Plan &p) {
    cudaError_t e = cudaDeviceGetAttribute(&p.sms, cudaDevAttrMultiProcessorCount, p.device);
    if (e != cudaSuccess) return e;
    const int t1 = p.n2 / 2, t2 = p.n1 / 4;
    switch (p.n1) {
        case 256: e = prepare_pass1<256>(t1, p.sms, p.grid1); break;
        case 512: e = prepare_pass1<512>(t1, p.sms, p.grid1); break;
        case 1024: e = prepare_pass1<1024>(t1, p.sms, p.grid1); break;
        case 2048: e = prepare_pass1<2048>(t1, p.sms, p.grid1); break;
        default: e = prepare_pass1<4096>(t1, p.sms, p.grid1); break;
    }
    if (e != cudaSuccess) return e;
    switch (p.n2) {
        case 256: return prepare_pass2<256>(t2, p.sms, &p.grid2);
        case 512: return prepare_pass2<512>(t2, p.sms, &p.grid2);
        case 1024: return prepare_pass2<1024>(t2, p.sms, &p.grid2);
        case 2048: return prepare_pass2<2048>(t2, p.sms, &p.grid2);
        default:
            if (k3_use_r64) {
                p.grid2 = t2 < p.sms ? t2 : p.sms;
                return cudaFuncSetAttribute(k3_pass2_r64, cudaFuncAttributeMaxDynamicSharedMemorySize, K3R_SMEM);
            }
            return prepare_pass2<4096>(t2, p.sms, &p.grid2);
    }
}

template <int N1, int KEEP> static cudaError_t launch_pass1k(const K3Plan &p, const CUtensorMap &a, float2 *mid, K3Tables tb,
                                                             cudaStream_t st, const float2 *in) {
    if (k3_chain)
        return launch_pdl(k3_pass1<N1, KEEP>, p.grid1[KEEP], N1 / 8, (size_t)(N1 * 24 + 16), st, a, mid, tb, p.n2 / 2, in);
    k3_pass1<N1, KEEP><<<p.grid1[KEEP], N1 / 8, N1 * 24 + 16, st>>>(a, mid, tb, p.n2 / 2, in);
    return cudaGetLastError();
}
template <int N1> static cudaError_t launch_pass1(const K3Plan &p, const CUtensorMap &a, float2 *mid, K3Tables tb, int keep,
                                                  cudaStream_t st, const float2 *in) {
    if (keep == 2) return launch_pass1k<N1, 2>(p, a, mid, tb, st, in);
    if (keep == 1) return launch_pass1k<N1, 1>(p, a, mid, tb, st, in);
    return launch_pass1k<N1, 0>(p, a, mid, tb, st, in);
}
template <int N2> static cudaError_t launch_pass2(const K3Plan &p, const float2 *mid, float2 *o, K3Tables tb,
                                                  int keep_from_tile, cudaStream_t st) {
    return launch_pdl(k3_pass2<N2>, p.grid2, N2 / 4, (size_t)(N2 * 40 + 16), st, mid, o, tb, p.n1 / 4, keep_from_tile);
}

}  // namespace kopek

using namespace kopek;

extern "C" kopek_status kopek_fft_large_c2c_f32(const float *d_in, float *d_out, uint64_t n, int device,
                                                void *cuda_stream) {
    if (!d_in || !d_out) return fail(KOPEK_ERR_INVALID, "NULL device pointer");
    if (d_in == d_out) return fail(KOPEK_ERR_INVALID, "in-place transform is not supported");
    if (n < (1u << 16) || n > (1u << 24) || (n & (n - 1)))
        return fail(KOPEK_ERR_UNSUPPORTED, "four-step transform: n=%llu must be a power of two in [2^16, 2^24]",
                    (unsigned long long)n);
    if (((uintptr_t)d_in | (uintptr_t)d_out) % 16) return fail(KOPEK_ERR_INVALID, "buffers must be 16-byte aligned");
    DeviceGuard g(device);
    if (!g.ok) return fail(KOPEK_ERR_CUDA, "cannot select device %d", device);
    K3Plan p;
    kopek_status st = k3_get_plan(n, device, &p);
    if (st != KOPEK_OK) return st;
    CUtensorMap in_map;
    // in:  [N1 rows][N2 complex], box 2 complex x 256 rows (pass 1: two-column tiles).  mid is tile-major
    // [k1/4][n2/2][k1%4][n2%2] (pass 1 stores 64-byte segments, pass 2 loads one contiguous block per tile); the
    // output is written in natural order by pass 2 in 32-byte runs.
    const uint32_t r1 = p.n1 < 256 ? p.n1 : 256;
    if (!make_map(&in_map, d_in, 2ull * p.n2, p.n1, 4, r1))
        return fail(KOPEK_ERR_CUDA, "cuTensorMapEncodeTiled failed (driver entry point missing or bad geometry)");
    K3Tables tb{p.tw1, p.tw2, p.hi, p.lo};
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    // workspace between the passes: this call's own, ordered on the caller's stream (allocation, both passes, release)
    cudaMemPool_t pool = nullptr;
    if ((st = k3_workspace_pool(device, &pool)) != KOPEK_OK) return st;
    float2 *mid = nullptr;
    cudaError_t e = cudaMallocFromPoolAsync(reinterpret_cast<void **>(&mid), n * sizeof(float2), pool, stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(KOPEK_ERR_NOMEM, "four-step workspace of %llu bytes: %s", (unsigned long long)(n * sizeof(float2)),
                    cudaGetErrorString(e));
    }
    struct Release {           // every exit path hands the block back in stream order
        float2 *ptr; cudaStream_t s;
        ~Release() { cudaFreeAsync(ptr, s); }
    } release{mid, stream};
    // L2 residency of the workspace: a `mid` of up to KOPEK_K3_KEEP_MB megabytes (default 72: n <= 2^23) is kept in
    // L2 as a whole (pass 1 stores it evict_last, pass 2 reads it there and drops the lines); a larger one keeps its
    // upper half of rows k1, which pass 2 reads first; 0 switches the keeping off (everything streams evict_first)
    static const long keep_mb = getenv("KOPEK_K3_KEEP_MB") ? atol(getenv("KOPEK_K3_KEEP_MB")) : 72;
    const int keep = keep_mb <= 0 ? 0 : n * sizeof(float2) <= ((uint64_t)keep_mb << 20) ? 2 : 1;
    const int tiles2 = p.n1 / 4;
    const int keep_from = (keep == 2 ? 0 : keep == 1 ? tiles2 / 2 : tiles2 + 1) * 4;     // in rows; the launches take tiles
    // L2 prefetch of a CTA's next pass-1 tile: on where measured to pay (n = 2^23); KOPEK_K3_PF1=0/1 forces it (tuning)
    static const int pf_env = getenv("KOPEK_K3_PF1") && *getenv("KOPEK_K3_PF1") ? atoi(getenv("KOPEK_K3_PF1")) : -1;
    const bool pf = pf_env >= 0 ? pf_env != 0 : n == (1ull << 23);
    const float2 *pf_in = pf ? reinterpret_cast<const float2 *>(d_in) : nullptr;
    e = cudaErrorInvalidValue;
    switch (p.n1) {
        case 256: e = launch_pass1<256>(p, in_map, mid, tb, keep, stream, pf_in); break;
        case 512: e = launch_pass1<512>(p, in_map, mid, tb, keep, stream, pf_in); break;
        case 1024: e = launch_pass1<1024>(p, in_map, mid, tb, keep, stream, pf_in); break;
        case 2048: e = launch_pass1<2048>(p, in_map, mid, tb, keep, stream, pf_in); break;
        case 4096: e = launch_pass1<4096>(p, in_map, mid, tb, keep, stream, pf_in); break;
    }
    if (e != cudaSuccess) return fail(KOPEK_ERR_CUDA, "four-step pass 1 (N1=%d): %s", p.n1, cudaGetErrorString(e));
    float2 *o = reinterpret_cast<float2 *>(d_out);
    switch (p.n2) {
        case 256: e = launch_pass2<256>(p, mid, o, tb, keep_from / 4, stream); break;
        case 512: e = launch_pass2<512>(p, mid, o, tb, keep_from / 4, stream); break;
        case 1024: e = launch_pass2<1024>(p, mid, o, tb, keep_from / 4, stream); break;
        case 2048: e = launch_pass2<2048>(p, mid, o, tb, keep_from / 4, stream); break;
        case 4096:
            if (k3_use_r64)
                e = launch_pdl(k3_pass2_r64, p.grid2, 256, (size_t)K3R_SMEM, stream, (const float2 *)mid, o,
                               (const float2 *)p.tw2r, p.n1 / 4, keep_from / 4);
            else
                e = launch_pass2<4096>(p, mid, o, tb, keep_from / 4, stream);
            break;
        default: e = cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return fail(KOPEK_ERR_CUDA, "four-step pass 2 (N2=%d): %s", p.n2, cudaGetErrorString(e));
    return KOPEK_OK;
}
