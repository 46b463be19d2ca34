// Host-visible launchers of the K1 kernel, one set per frame size (k1_inst.cu).
#pragma once
#include "fft_batched.cuh"

namespace kopek {

struct K1Info {
    int block;   // threads per CTA
    int fpb;     // frames per CTA per loop iteration
    int smem;    // dynamic shared memory bytes
};

#define KOPEK_K1_DECL(N_)                                                                               \
    cudaError_t k1_launch_##N_(const K1Params &p, int output, bool win, bool full, bool vec, int grid,  \
                               cudaStream_t stream);                                                    \
    int k1_occupancy_##N_(int output, bool win);                                                        \
    K1Info k1_info_##N_();
KOPEK_K1_DECL(256)
KOPEK_K1_DECL(512)
KOPEK_K1_DECL(1024)
KOPEK_K1_DECL(2048)
KOPEK_K1_DECL(4096)
#undef KOPEK_K1_DECL

}  // namespace kopek
