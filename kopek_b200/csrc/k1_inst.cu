// Instantiations of the K1 batched STFT kernel for one frame size.  Compiled
// once per size with -DKOPEK_K1_N=<256|512|1024|2048|4096> (see kopek_b200/build.py)
// so the five sizes build in parallel.
#include "k1_launch.cuh"

#ifndef KOPEK_K1_N
#error "compile with -DKOPEK_K1_N=<frame size>"
#endif

namespace kopek {

template <int OUT, bool WIN>
static cudaError_t launch_one(const K1Params &p, bool full, bool vec, int grid, cudaStream_t stream) {
    using G = K1Geom<KOPEK_K1_N>;
    auto kern = k1_stft_kernel<KOPEK_K1_N, OUT, WIN>;
    static bool attr_done = false;   // benign race: idempotent attribute
    if (!attr_done && G::SMEM > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
        if (e != cudaSuccess) return e;
    }
    attr_done = true;
    K1Params q = p;
    q.full = full ? 1 : 0;
    q.vec = vec ? 1 : 0;
    kern<<<grid, G::BLOCK, G::SMEM, stream>>>(q);
    return cudaGetLastError();
}

template <int OUT, bool WIN> static int occupancy_one() {
    using G = K1Geom<KOPEK_K1_N>;
    auto kern = k1_stft_kernel<KOPEK_K1_N, OUT, WIN>;
    if (G::SMEM > 48 * 1024)
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, G::BLOCK, G::SMEM) != cudaSuccess) return 0;
    return nb;
}

#define KOPEK_DISPATCH(FN_, ...)                                                      \
    switch (output * 2 + (win ? 1 : 0)) {                                             \
        case 0: return FN_<OUT_COMPLEX, false>(__VA_ARGS__);                          \
        case 1: return FN_<OUT_COMPLEX, true>(__VA_ARGS__);                           \
        case 2: return FN_<OUT_MAG, false>(__VA_ARGS__);                              \
        case 3: return FN_<OUT_MAG, true>(__VA_ARGS__);                               \
        case 4: return FN_<OUT_POWER, false>(__VA_ARGS__);                            \
        case 5: return FN_<OUT_POWER, true>(__VA_ARGS__);                             \
        case 6: return FN_<OUT_DB, false>(__VA_ARGS__);                               \
        default: return FN_<OUT_DB, true>(__VA_ARGS__);                               \
    }

#define KOPEK_CAT2(a, b) a##b
#define KOPEK_CAT(a, b) KOPEK_CAT2(a, b)

cudaError_t KOPEK_CAT(k1_launch_, KOPEK_K1_N)(const K1Params &p, int output, bool win, bool full, bool vec,
                                              int grid, cudaStream_t stream) {
    KOPEK_DISPATCH(launch_one, p, full, vec, grid, stream)
}
int KOPEK_CAT(k1_occupancy_, KOPEK_K1_N)(int output, bool win) {
    KOPEK_DISPATCH(occupancy_one)
}
K1Info KOPEK_CAT(k1_info_, KOPEK_K1_N)() {
    using G = K1Geom<KOPEK_K1_N>;
    return K1Info{G::BLOCK, G::FPB, (int)G::SMEM};
}

}  // namespace kopek
