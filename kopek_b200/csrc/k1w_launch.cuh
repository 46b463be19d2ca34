// Host-visible launchers of the lane-group kernels (k1w_inst.cu, k1m_inst.cu).
// occ_out != nullptr: no launch, returns the resident warps (teams) per SM of that instantiation.
#pragma once
#include "lane_common.cuh"

namespace kopek {

// 1024-point two-exchange kernel (fft_warp1024b.cuh): 4 warps per CTA, one frame per warp
constexpr int K1B_TABLE_F4 = (512 + 480 + 16 + 32) / 2;
void k1b_build_tables(const float *half_window, float4 *out /* K1B_TABLE_F4 */);
cudaError_t k1b_launch(const K1WParams &p, int output, bool win, bool bulk, int grid, cudaStream_t stream,
                       int *occ_out);
// same kernel writing all 1024 bins per row (the conjugate mirror stored beside the computed half)
cudaError_t k1b_launch_full(const K1WParams &p, int output, bool win, int grid, cudaStream_t stream);
// same kernel reading interleaved stereo i16 frames (p.x, p.hop in frames, p.pcm_shift selects the channel)
cudaError_t k1b_launch_pcm(const K1WParams &p, int output, bool win, int grid, cudaStream_t stream);

// 256-point kernel, eight lanes per frame (fft_warp256.cuh): 4 warps per CTA, 4 frames per warp
constexpr int K1W256_TABLE_F4 = 64 + 56 + 8;
void k1w256_build_tables(const float *half_window, float4 *out /* K1W256_TABLE_F4 */);
// mode: bit 0 = FULL rows, bit 1 = 32-bit loads (odd hop / unaligned base)
cudaError_t k1w256_launch(const K1WParams &p, int output, bool win, int mode, int grid, cudaStream_t stream, int *occ_out);

// 512-point kernel, sixteen lanes per frame (fft_warp512.cuh): 4 warps per CTA, 2 frames per warp
constexpr int K1W512_TABLE_F4 = (256 + 240 + 16) / 2;
void k1w512_build_tables(const float *half_window, float4 *out /* K1W512_TABLE_F4 */);
cudaError_t k1w512_launch(const K1WParams &p, int output, bool win, int mode, int grid, cudaStream_t stream, int *occ_out);

// team-per-frame kernels (fft_team.cuh): n = 2048 (64 lanes) / 4096 (128 lanes)
int k1m_table_f4(int n);
// wink: 0 rectangular, 2 cosine-sum window w[n] = cs_a - cs_b cos(2 pi n / n) on the fly, 3 window table in shared
// memory (half_window: custom windows)
void k1m_build_tables(int n, const float *half_window, double cs_a, double cs_b, float4 *out /* k1m_table_f4(n) */);
// mode: 0 HALF rows + 64-bit loads, 1 FULL rows, 2 HALF + 32-bit loads (odd hop / unaligned base), 3 FULL + 32-bit
// loads
cudaError_t k1m_launch(int n, const K1WParams &p, int output, int wink, int mode, int grid, cudaStream_t stream,
                       int *occ_out /* teams per SM */);

}  // namespace kopek
