// Decoder-side callers of the path, on i16 PCM frames as kopek::decoder::decode returns them
// (src/decoder.rs:1-31: Vec<[i16; 2]>):
//   kopek_pcm16_to_f32  -- the marshalling of examples/graph/player.rs:56-59 (frame[0] as f64 / i16::MAX) into the f32
//                          sample buffer the STFT plans transform; bit-exact (f64 division, one narrowing)
//   kopek_pcm16_scale_f32 -- the marshalling of examples/graph/player.rs:81-102 (load_track_at_path): every i16 value of
//                          the decoded frames, channels interleaved, becomes volume_factor * (value as f32); bit-exact
//   kopek_f32_downmix   -- the capture callback of examples/pong/audio.rs:57-70: every interleaved f32 frame becomes
//                          ((0.0 + f[0]) + f[1]) / channels (the first TWO channels summed, divided by ALL); bit-exact
//   kopek_detect_bpm    -- the energy pass of detect_bpm (src/decoder.rs:39-71); bit-exact: the per-frame terms are
//                          computed in parallel (no contraction), and every f32 sum of the reference -- a sequential
//                          left fold -- is added in the reference's order by ONE thread per 1024-frame block (43 per
//                          second) and one thread per one-second average; the parallelism is across blocks and seconds.
#include <algorithm>
#include <mutex>

#include "common.cuh"

namespace kopek {

__global__ void __launch_bounds__(256) pcm16_to_f32_kernel(const int16_t *__restrict__ pcm, uint64_t n_frames,
                                                           uint32_t channels, uint32_t channel, double divisor,
                                                           float *__restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_frames; i += stride)
        out[i] = __double2float_rn(__ddiv_rn((double)pcm[i * channels + channel], divisor));
}
// stereo, 16-byte aligned: four frames per thread (one 128-bit load, one 128-bit store)
__global__ void __launch_bounds__(256) pcm16x2_to_f32_kernel(const int4 *__restrict__ pcm, uint64_t n_quads, uint32_t channel,
                                                             double divisor, float4 *__restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_quads; i += stride) {
        const int4 v = __ldg(pcm + i);
        const int w[4] = {v.x, v.y, v.z, v.w};
        float r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int s = channel ? (w[k] >> 16) : (int)(short)(w[k] & 0xffff);
            r[k] = __double2float_rn(__ddiv_rn((double)s, divisor));
        }
        out[i] = make_float4(r[0], r[1], r[2], r[3]);
    }
}

// load_track_at_path (examples/graph/player.rs:87-97): [volume_factor * frame[0] as f32, volume_factor * frame[1] as f32]
// flattened -- one f32 product per i16 value.  Eight values per thread: one 128-bit load, two 128-bit stores.
__device__ __forceinline__ void pcm16_scale_octet(const int4 v, float factor, float4 *out) {
    const int w[4] = {v.x, v.y, v.z, v.w};
    float r[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        r[2 * k] = __fmul_rn(factor, (float)(short)(w[k] & 0xffff));
        r[2 * k + 1] = __fmul_rn(factor, (float)(w[k] >> 16));
    }
    __stcs(out, make_float4(r[0], r[1], r[2], r[3]));          // the track is written once and not read back here
    __stcs(out + 1, make_float4(r[4], r[5], r[6], r[7]));
}
__global__ void __launch_bounds__(256) pcm16_scale_f32_kernel(const int4 *__restrict__ pcm, uint64_t n_octets, float factor,
                                                              float4 *__restrict__ out) {
    // four independent 128-bit loads in flight per thread and iteration (a warp covers 4 x 512 contiguous bytes)
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 4;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x * 4 + threadIdx.x; i0 < n_octets; i0 += stride) {
        int4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i0 + u * 256 < n_octets) v[u] = __ldcs(pcm + i0 + u * 256);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i0 + u * 256 < n_octets) pcm16_scale_octet(v[u], factor, out + 2 * (i0 + u * 256));
    }
}
__global__ void __launch_bounds__(256) pcm16_scale_f32_tail_kernel(const int16_t *__restrict__ pcm, uint64_t n, float factor,
                                                                   float *__restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = __fmul_rn(factor, (float)pcm[i]);
}

// examples/pong/audio.rs:60-66: average = 0.0; for sample in frame.iter().take(2) { average += sample }; average /= len
// (0.0 + x is x except that it turns -0.0 into +0.0: kept).  Stereo: one 64-bit load per frame.
__global__ void __launch_bounds__(256) f32_downmix_kernel(const float *__restrict__ frames, uint64_t n_frames, uint32_t channels,
                                                          float *__restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const float len = (float)channels;
    const bool vec2 = channels == 2 && ((uintptr_t)frames % 8) == 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_frames; i += stride) {
        float a, b = 0.0f;
        if (vec2) {
            const float2 f = __ldcs(reinterpret_cast<const float2 *>(frames) + i);
            a = f.x, b = f.y;
        } else {
            a = __ldg(frames + i * channels);
            if (channels > 1) b = __ldg(frames + i * channels + 1);
        }
        float avg = __fadd_rn(0.0f, a);
        if (channels > 1) avg = __fadd_rn(avg, b);
        out[i] = __fdiv_rn(avg, len);
    }
}

// f[0].powf(2.0) + f[1].powf(2.0) of one frame (src/decoder.rs:52,61), f = s as f32 / i16::MAX as f32
__device__ __forceinline__ float frame_energy(int frame) {
    const float a = __fdiv_rn((float)(short)(frame & 0xffff), 32767.0f), b = __fdiv_rn((float)(frame >> 16), 32767.0f);
    return __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
}

// step 1, fully parallel: the per-frame term of both sums, four frames per thread (128-bit load, 128-bit store)
__global__ void __launch_bounds__(256) bpm_frame_energy_kernel(const int4 *__restrict__ frames, uint64_t n_quads,
                                                               float4 *__restrict__ term) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_quads; i += stride) {
        const int4 v = __ldg(frames + i);
        term[i] = make_float4(frame_energy(v.x), frame_energy(v.y), frame_energy(v.z), frame_energy(v.w));
    }
}

// step 2, the sequential folds (Iterator::sum is a left fold from 0.0): ONE WARP per sum.  The lanes load 32 x 4 terms
// coalesced (the next step's load is issued before the current one is folded), then every lane adds the 128 values in
// index order, fetching them by shuffle: all lanes carry the same running sum, so the chain of 4-cycle adds -- not
// the memory latency -- sets the pace (a single thread with eight loads in flight was 7 x slower on a 3 s clip).
// warp w < n_blocks: block (i, j) = (w / 43, w % 43), terms i*44100 + j*1024 .. +1024         (decoder.rs:57-62)
// warp w >= n_blocks: second i = w - n_blocks, terms i*44100 .. +44100, scaled by 1024/44100   (decoder.rs:51-54)
__global__ void __launch_bounds__(128) bpm_fold_kernel(const float *__restrict__ term, uint64_t n_blocks, uint64_t n_seconds,
                                                       float *__restrict__ block_e, float *__restrict__ avg_e) {
    const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_blocks + n_seconds) return;
    const bool is_block = w < n_blocks;
    const uint64_t sec = is_block ? w / 43 : w - n_blocks;
    const uint64_t start = sec * 44100 + (is_block ? (w % 43) * 1024 : 0);
    const uint32_t quads = is_block ? 256u : 11025u;
    const float4 *p = reinterpret_cast<const float4 *>(term + start);    // 16-byte aligned: 44100 and 1024 are multiples of 4
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    float e = 0.0f;
    float4 nxt = lane < (int)quads ? __ldg(p + lane) : zero;
    for (uint32_t k = 0; k < quads; k += 32) {
        const float4 cur = nxt;
        if (k + 32 < quads) nxt = k + 32 + lane < quads ? __ldg(p + k + 32 + lane) : zero;
        const int valid = quads - k < 32u ? (int)(quads - k) : 32;
        if (valid == 32) {
#pragma unroll
            for (int src = 0; src < 32; ++src) {
                e = __fadd_rn(e, __shfl_sync(0xffffffffu, cur.x, src));
                e = __fadd_rn(e, __shfl_sync(0xffffffffu, cur.y, src));
                e = __fadd_rn(e, __shfl_sync(0xffffffffu, cur.z, src));
                e = __fadd_rn(e, __shfl_sync(0xffffffffu, cur.w, src));
            }
        } else {
            for (int src = 0; src < valid; ++src) {
                e = __fadd_rn(e, __shfl_sync(0xffffffffu, cur.x, src));
                e = __fadd_rn(e, __shfl_sync(0xffffffffu, cur.y, src));
                e = __fadd_rn(e, __shfl_sync(0xffffffffu, cur.z, src));
                e = __fadd_rn(e, __shfl_sync(0xffffffffu, cur.w, src));
            }
        }
    }
    if (lane == 0) {
        if (is_block) block_e[w] = e;
        else avg_e[sec] = __fmul_rn(e, 1024.0f / 44100.0f);
    }
}

__global__ void __launch_bounds__(256) bpm_count_kernel(const float *__restrict__ block_e, const float *__restrict__ avg_e,
                                                        uint64_t n_blocks, unsigned int *beats) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool beat = t < n_blocks && block_e[t] > __fmul_rn(5.5f, avg_e[t / 43]);            // decoder.rs:64 (C = 5.5)
    const unsigned int n = __popc(__ballot_sync(0xffffffffu, beat));
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(beats, n);
}

// one stream-ordered memory pool per device, release threshold "never": scratch stays mapped between calls
static cudaMemPool_t scratch_pool(int device) {
    static std::mutex mu;
    static cudaMemPool_t pools[64] = {};
    if (device < 0 || device >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    if (!pools[device]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaMemPool_t pool = nullptr;
        if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) return nullptr;
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        pools[device] = pool;
    }
    return pools[device];
}

}  // namespace kopek

using namespace kopek;

extern "C" kopek_status kopek_pcm16_to_f32(const int16_t *d_pcm, uint64_t n_frames, uint32_t channels, uint32_t channel,
                                           double divisor, float *d_out, int device, void *cuda_stream) {
    if (n_frames == 0) return KOPEK_OK;
    if (!d_pcm || !d_out) return fail(KOPEK_ERR_INVALID, "NULL device pointer");
    if (channels == 0 || channel >= channels) return fail(KOPEK_ERR_INVALID, "channel %u of %u", channel, channels);
    if (!(divisor != 0.0)) return fail(KOPEK_ERR_INVALID, "divisor must be non-zero");
    if ((uintptr_t)d_pcm % 2 || (uintptr_t)d_out % 4) return fail(KOPEK_ERR_INVALID, "misaligned device pointer");
    DeviceGuard g(device);
    if (!g.ok) return fail(KOPEK_ERR_CUDA, "cannot select device %d", device);
    int sms = 0;
    KOPEK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    uint64_t done = 0;
    if (channels == 2 && (uintptr_t)d_pcm % 16 == 0 && (uintptr_t)d_out % 16 == 0 && n_frames >= 4) {
        const uint64_t quads = n_frames / 4;
        const unsigned grid = (unsigned)std::min<uint64_t>((quads + 255) / 256, (uint64_t)sms * 8);
        pcm16x2_to_f32_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const int4 *>(d_pcm), quads, channel, divisor,
                                                    reinterpret_cast<float4 *>(d_out));
        done = quads * 4;
    }
    if (done < n_frames) {
        const uint64_t rest = n_frames - done;
        const unsigned grid = (unsigned)std::min<uint64_t>((rest + 255) / 256, (uint64_t)sms * 8);
        pcm16_to_f32_kernel<<<grid, 256, 0, st>>>(d_pcm + done * channels, rest, channels, channel, divisor, d_out + done);
    }
    KOPEK_CUDA(cudaGetLastError());
    return KOPEK_OK;
}

extern "C" kopek_status kopek_pcm16_scale_f32(const int16_t *d_pcm, uint64_t n_values, float factor, float *d_out, int device,
                                              void *cuda_stream) {
    if (n_values == 0) return KOPEK_OK;
    if (!d_pcm || !d_out) return fail(KOPEK_ERR_INVALID, "NULL device pointer");
    if ((uintptr_t)d_pcm % 2 || (uintptr_t)d_out % 4) return fail(KOPEK_ERR_INVALID, "misaligned device pointer");
    DeviceGuard g(device);
    if (!g.ok) return fail(KOPEK_ERR_CUDA, "cannot select device %d", device);
    int sms = 0;
    KOPEK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    uint64_t done = 0;
    if ((uintptr_t)d_pcm % 16 == 0 && (uintptr_t)d_out % 16 == 0 && n_values >= 8) {
        const uint64_t octets = n_values / 8;
        const unsigned grid = (unsigned)std::min<uint64_t>((octets + 1023) / 1024, (uint64_t)sms * 8);
        pcm16_scale_f32_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const int4 *>(d_pcm), octets, factor,
                                                     reinterpret_cast<float4 *>(d_out));
        done = octets * 8;
    }
    if (done < n_values) {
        const uint64_t rest = n_values - done;
        const unsigned grid = (unsigned)std::min<uint64_t>((rest + 255) / 256, (uint64_t)sms * 8);
        pcm16_scale_f32_tail_kernel<<<grid, 256, 0, st>>>(d_pcm + done, rest, factor, d_out + done);
    }
    KOPEK_CUDA(cudaGetLastError());
    return KOPEK_OK;
}

extern "C" kopek_status kopek_f32_downmix(const float *d_frames, uint64_t n_frames, uint32_t channels, float *d_out, int device,
                                          void *cuda_stream) {
    if (n_frames == 0) return KOPEK_OK;
    if (!d_frames || !d_out) return fail(KOPEK_ERR_INVALID, "NULL device pointer");
    if (channels == 0) return fail(KOPEK_ERR_INVALID, "channels must be >= 1");
    if ((uintptr_t)d_frames % 4 || (uintptr_t)d_out % 4) return fail(KOPEK_ERR_INVALID, "misaligned device pointer");
    DeviceGuard g(device);
    if (!g.ok) return fail(KOPEK_ERR_CUDA, "cannot select device %d", device);
    int sms = 0;
    KOPEK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const unsigned grid = (unsigned)std::min<uint64_t>((n_frames + 255) / 256, (uint64_t)sms * 16);
    f32_downmix_kernel<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(d_frames, n_frames, channels, d_out);
    KOPEK_CUDA(cudaGetLastError());
    return KOPEK_OK;
}

extern "C" kopek_status kopek_detect_bpm(const int16_t *d_frames, uint64_t n_frames, float sample_rate, uint32_t *bpm_out,
                                         uint32_t *beats_out, float *d_block_e, float *d_avg_e, int device,
                                         void *cuda_stream) {
    if (!d_frames || !bpm_out) return fail(KOPEK_ERR_INVALID, "NULL argument");
    if (!(sample_rate > 0.0f)) return fail(KOPEK_ERR_INVALID, "sample_rate must be positive");
    if ((uintptr_t)d_frames % 16) return fail(KOPEK_ERR_INVALID, "frames must be 16-byte aligned");
    // duration_in_seconds(...) as usize (decoder.rs:34-37,46); the Rust original panics for 0 (60 / 0) and when the
    // hard-coded 44100-frame seconds run past the slice: both are errors here
    const uint64_t duration = (uint64_t)((float)n_frames / sample_rate);
    if (duration == 0) return fail(KOPEK_ERR_INVALID, "detect_bpm: less than one second of audio (the reference divides by zero)");
    if (duration * 44100 > n_frames)
        return fail(KOPEK_ERR_INVALID, "detect_bpm: %llu seconds x 44100 frames exceed the %llu frames given (the reference "
                    "indexes out of bounds)", (unsigned long long)duration, (unsigned long long)n_frames);
    DeviceGuard g(device);
    if (!g.ok) return fail(KOPEK_ERR_CUDA, "cannot select device %d", device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int sms = 0;
    KOPEK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const uint64_t n_blocks = duration * 43, n_used = duration * 44100;
    // stream-ordered scratch from the library's own pool (kept mapped between calls: no device synchronisation, no
    // page mapping per call): per-frame terms | block_e | avg_e | beats
    cudaMemPool_t pool = scratch_pool(device);
    if (!pool) return fail(KOPEK_ERR_CUDA, "detect_bpm: cannot create the scratch pool: %s", cudaGetErrorString(cudaGetLastError()));
    float *scratch = nullptr;
    cudaError_t e = cudaMallocFromPoolAsync(&scratch, (n_used + n_blocks + duration) * sizeof(float) + 16, pool, st);
    if (e != cudaSuccess) return fail(KOPEK_ERR_NOMEM, "detect_bpm scratch: %s", cudaGetErrorString(e));
    float *term = scratch;
    float *block_e = d_block_e ? d_block_e : scratch + n_used;
    float *avg_e = d_avg_e ? d_avg_e : scratch + n_used + n_blocks;
    unsigned int *d_beats = reinterpret_cast<unsigned int *>(scratch + n_used + n_blocks + duration);
    unsigned int beats = 0;
    const uint64_t threads = n_blocks + duration, quads = n_used / 4;
    if ((e = cudaMemsetAsync(d_beats, 0, sizeof(unsigned int), st)) == cudaSuccess) {
        const unsigned g1 = (unsigned)std::min<uint64_t>((quads + 255) / 256, (uint64_t)sms * 16);
        bpm_frame_energy_kernel<<<g1, 256, 0, st>>>(reinterpret_cast<const int4 *>(d_frames), quads,
                                                    reinterpret_cast<float4 *>(term));
        bpm_fold_kernel<<<(unsigned)((threads + 3) / 4), 128, 0, st>>>(term, n_blocks, duration, block_e, avg_e);   // one warp per sum
        bpm_count_kernel<<<(unsigned)((n_blocks + 255) / 256), 256, 0, st>>>(block_e, avg_e, n_blocks, d_beats);
        if ((e = cudaGetLastError()) == cudaSuccess)
            e = cudaMemcpyAsync(&beats, d_beats, sizeof(beats), cudaMemcpyDeviceToHost, st);
    }
    cudaFreeAsync(scratch, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(KOPEK_ERR_CUDA, "detect_bpm: %s", cudaGetErrorString(e));
    if (beats_out) *beats_out = beats;
    *bpm_out = beats * (60u / (uint32_t)duration);                                          // decoder.rs:70, u32 division
    return KOPEK_OK;
}

// host-buffer form: the exact shape of kopek::decoder::detect_bpm(frames: Vec<[i16; 2]>, sample_rate: f32) -> u32
extern "C" kopek_status kopek_detect_bpm_host(const int16_t *h_frames, uint64_t n_frames, float sample_rate,
                                              uint32_t *bpm_out, uint32_t *beats_out, int device) {
    if (!h_frames || !bpm_out) return fail(KOPEK_ERR_INVALID, "NULL argument");
    if (!(sample_rate > 0.0f)) return fail(KOPEK_ERR_INVALID, "sample_rate must be positive");
    const uint64_t duration = (uint64_t)((float)n_frames / sample_rate);
    if (duration == 0) return fail(KOPEK_ERR_INVALID, "detect_bpm: less than one second of audio (the reference divides by zero)");
    if (duration * 44100 > n_frames)
        return fail(KOPEK_ERR_INVALID, "detect_bpm: %llu seconds x 44100 frames exceed the %llu frames given (the reference "
                    "indexes out of bounds)", (unsigned long long)duration, (unsigned long long)n_frames);
    DeviceGuard g(device);
    if (!g.ok) return fail(KOPEK_ERR_CUDA, "cannot select device %d", device);
    cudaMemPool_t pool = scratch_pool(device);
    if (!pool) return fail(KOPEK_ERR_CUDA, "detect_bpm: cannot create the scratch pool: %s", cudaGetErrorString(cudaGetLastError()));
    // only the frames the algorithm reads (duration * 44100) cross PCIe
    const uint64_t used = duration * 44100;
    cudaStream_t st = nullptr;
    KOPEK_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int16_t *d_frames = nullptr;
    cudaError_t e = cudaMallocFromPoolAsync(&d_frames, used * 4, pool, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_frames, h_frames, used * 4, cudaMemcpyHostToDevice, st);
    kopek_status rc = KOPEK_OK;
    if (e != cudaSuccess) rc = fail(KOPEK_ERR_CUDA, "detect_bpm: staging the frames failed: %s", cudaGetErrorString(e));
    // same (n_frames, sample_rate) as given, so the duration is computed from the same f32 quotient; the device
    // routine reads the first duration * 44100 frames only
    else rc = kopek_detect_bpm(d_frames, n_frames, sample_rate, bpm_out, beats_out, nullptr, nullptr, device, st);
    if (d_frames) cudaFreeAsync(d_frames, st);
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    return rc;
}
