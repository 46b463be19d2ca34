// Shared pieces of the lane-group kernels (fft_warp256/512/1024b.cuh, fft_team.cuh) and of the four-step passes
// (fft_large.cu): the launch parameter block, TMA (bulk async copy) + mbarrier primitives, streaming loads and the
// compile-time twiddle multiplies.  (The three-pass 1024-point kernel these came from was superseded in round 1 by
// fft_warp1024b.cuh; it is kept for reference under tools/legacy/.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fft_batched.cuh"

namespace kopek {

struct K1WParams {
    const float *x;          // samples
    void *out;               // spectra
    const float4 *tables;    // the kernel's own table block (window | twiddles | post twiddles), see *_build_tables
    uint64_t n_frames;
    uint64_t hop;
    uint64_t pitch;
    float band_scale;        // OUT_BANDS_*: display scale applied to the band means
    uint32_t pcm_shift;      // PCM input (fft_warp1024b.cuh): 16 = channel 0 (low half of a stereo frame), 0 = channel 1
    uint32_t vec;            // host side only: 1 = every frame starts 8-byte aligned (base % 8 == 0, hop even), the
                             // launcher picks the 64-bit-load instantiation; 0 = the two-32-bit-loads-per-pair one
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

}  // namespace kopek
