// K4 -- on-device signal sources: the reference's Oscillator (src/oscillator.rs:34-91) and Noise
// (src/noise.rs:15-21) generating frames directly in HBM, so that a spectrum job needs no host-side synthesis and
// no H2D copy.  One thread owns one frame and runs the reference's f32 phase-accumulator recurrence verbatim
//     value = wave(phase);  phase = (phase + 2 pi f / sr) % (2 pi)      (sine / fake_sine / sawtooth)
//     value = wave(phase);  phase = fract(phase + f / sr)              (square / triangle)
// starting from phase 0 (Oscillator::new), so the PHASE sequence is bit-identical to the CPU (IEEE add, exact fmod);
// only sinf differs from glibc by <= 2 ulp.  Frames are independent, which is what makes the sequential recurrence
// parallel.  Noise is counter-based (one Philox-like hash per sample) because the reference's is an unseeded
// ThreadRng: only the distributions (U[-1,1), N(0,1)) can be matched.
#include <math.h>

#include "common.cuh"

namespace kopek {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {       // lowbias32 integer hash
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ float u01(uint64_t seed, uint64_t idx, uint32_t stream) {
    uint32_t h = mix32((uint32_t)idx ^ mix32((uint32_t)(idx >> 32) + 0x9e3779b9U * stream) ^ mix32((uint32_t)seed) ^
                       (uint32_t)(seed >> 32));
    return (float)(h >> 8) * (1.0f / 16777216.0f);             // [0, 1), 24 bits like rand's f32
}

struct SynthParams {
    float *out;
    uint64_t n_frames;
    uint32_t n;                 // samples per frame
    uint64_t stride;            // samples between frame starts in `out`
    float sample_rate, freq0, freq_step;
    uint32_t freq_mod;          // frame f runs at freq0 + freq_step * (f % freq_mod)
    uint32_t wave;              // 0 sine, 1 fake_sine, 2 sawtooth, 3 square, 4 triangle, 5 silence
    float duty;
    uint32_t noise;             // 0 none, 1 uniform [-1,1) (rand_noise), 2 N(0,1) (white_noise)
    float noise_gain;
    uint64_t seed;
    uint64_t first_frame;       // global index of frame 0 (sharding keeps the sequence identical)
};

__global__ void __launch_bounds__(128) synth_frames_kernel(const SynthParams p) {
    const uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= p.n_frames) return;
    const uint64_t gf = p.first_frame + f;
    const float freq = p.freq0 + p.freq_step * (float)(p.freq_mod ? gf % p.freq_mod : 0);
    const float pi = 3.14159265358979323846f;
    float phase = 0.0f;
    float *o = p.out + f * p.stride;
    for (uint32_t i = 0; i < p.n; ++i) {
        float value;
        // Oscillator::run returns 0.0 below f32::EPSILON (src/oscillator.rs:34-37)
        if (p.wave == 5 || freq < 1.1920929e-7f) {
            value = 0.0f;
        } else if (p.wave == 0) {
            value = sinf(phase);
            phase = fmodf(__fadd_rn(phase, __fdiv_rn(__fmul_rn(__fmul_rn(2.0f, pi), freq), p.sample_rate)), __fmul_rn(2.0f, pi));
        } else if (p.wave == 1) {
            float x = __fsub_rn(__fdiv_rn(phase, pi), 1.0f);
            value = __fmul_rn(__fmul_rn(4.0f, x), __fsub_rn(1.0f, fabsf(x)));
            phase = fmodf(__fadd_rn(phase, __fdiv_rn(__fmul_rn(__fmul_rn(2.0f, pi), freq), p.sample_rate)), __fmul_rn(2.0f, pi));
        } else if (p.wave == 2) {
            value = __fsub_rn(__fmul_rn(2.0f, __fdiv_rn(phase, __fmul_rn(2.0f, pi))), 1.0f);
            phase = fmodf(__fadd_rn(phase, __fdiv_rn(__fmul_rn(__fmul_rn(2.0f, pi), freq), p.sample_rate)), __fmul_rn(2.0f, pi));
        } else if (p.wave == 3) {
            float d = fminf(fmaxf(p.duty, 0.0f), 1.0f);
            value = phase < d ? 1.0f : -1.0f;
            float q = __fadd_rn(phase, __fdiv_rn(freq, p.sample_rate));
            phase = __fsub_rn(q, truncf(q));
        } else {
            value = __fsub_rn(1.0f, __fmul_rn(4.0f, fabsf(__fsub_rn(phase, 0.5f))));
            float q = __fadd_rn(phase, __fdiv_rn(freq, p.sample_rate));
            phase = __fsub_rn(q, truncf(q));
        }
        if (p.noise == 1) {
            value += p.noise_gain * (u01(p.seed, gf * p.n + i, 0) * 2.0f - 1.0f);       // noise.rs:15-17
        } else if (p.noise == 2) {                                                      // Box-Muller N(0,1)
            float u1 = 1.0f - u01(p.seed, gf * p.n + i, 1), u2 = u01(p.seed, gf * p.n + i, 2);
            value += p.noise_gain * sqrtf(-2.0f * logf(u1)) * cosf(2.0f * pi * u2);
        }
        o[i] = value;
    }
}

}  // namespace kopek

using namespace kopek;

extern "C" kopek_status kopek_synth_frames(float *d_out, uint64_t n_frames, uint32_t n, uint64_t stride,
                                           float sample_rate, uint32_t wave, float duty, float freq0, float freq_step,
                                           uint32_t freq_mod, uint32_t noise, float noise_gain, uint64_t seed,
                                           uint64_t first_frame, int device, void *cuda_stream) {
    if (n_frames == 0 || n == 0) return KOPEK_OK;
    if (!d_out) return fail(KOPEK_ERR_INVALID, "d_out is NULL");
    if (wave > 5 || noise > 2) return fail(KOPEK_ERR_INVALID, "unknown wave %u / noise %u", wave, noise);
    if (stride < n) return fail(KOPEK_ERR_INVALID, "stride %llu < n %u", (unsigned long long)stride, n);
    if (!(sample_rate > 0.0f)) return fail(KOPEK_ERR_INVALID, "sample_rate must be positive");
    DeviceGuard g(device);
    if (!g.ok) return fail(KOPEK_ERR_CUDA, "cannot select device %d", device);
    SynthParams p{d_out, n_frames, n, stride, sample_rate, freq0, freq_step, freq_mod, wave, duty, noise, noise_gain,
                  seed, first_frame};
    const uint64_t blocks = (n_frames + 127) / 128;
    if (blocks > 0x7fffffffULL) return fail(KOPEK_ERR_UNSUPPORTED, "too many frames for one launch");
    synth_frames_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)cuda_stream>>>(p);
    KOPEK_CUDA(cudaGetLastError());
    return KOPEK_OK;
}
