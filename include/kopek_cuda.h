/*
 * kopek_cuda.h -- C ABI of libkopek_cuda.so, the B200 (sm_100a) implementation
 * of kopek's FFT frequency-analysis path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types,
 * never throws or unwinds.  Every entry point returns a kopek_status (0 = OK,
 * < 0 = error) unless stated; the message of the last failure on the calling
 * thread is available from kopek_last_error().  There is NO CPU fallback: if no
 * CUDA device is usable the calls fail with KOPEK_ERR_CUDA.
 *
 * Reference interfaces replaced (paths relative to /root/reference):
 *   kopek_fft_c2c_f64   <- pub fn fft(&[Complex<f64>]) -> Vec<Complex<f64>>   src/fft.rs:6-41
 *   kopek_show(_format) <- pub fn show(label: &str, buf: &[Complex<f64>])       src/fft.rs:43-52
 *   kopek_stft_*        <- the per-UI-frame caller blocks that marshal f32 samples,
 *                          call fft and take magnitudes:
 *                            examples/beep/view.rs:140-158   (norm(), dB in comment :155)
 *                            examples/graph/player.rs:56-61 + examples/graph/utils.rs:47-63
 *                            examples/pong/main.rs:192-196  + examples/pong/utils.rs:36-47
 * Rust binding: rust/kopek-cuda-sys/src/lib.rs; see INTEGRATION.md.
 */
#ifndef KOPEK_CUDA_H
#define KOPEK_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define KOPEK_API __attribute__((visibility("default")))
#else
#define KOPEK_API
#endif

typedef int32_t kopek_status;
enum {
    KOPEK_OK = 0,
    KOPEK_ERR_INVALID = -1,     /* bad argument (null pointer, size, enum value, alignment) */
    KOPEK_ERR_CUDA = -2,        /* a CUDA runtime call failed; see kopek_last_error()       */
    KOPEK_ERR_NOMEM = -3,       /* host or device allocation failed                         */
    KOPEK_ERR_UNSUPPORTED = -4  /* valid request this build has no kernel for               */
};

enum { KOPEK_WINDOW_RECT = 0, KOPEK_WINDOW_HANN = 1, KOPEK_WINDOW_HAMMING = 2, KOPEK_WINDOW_CUSTOM = 3 };
/* What the fused epilogue writes per bin.  COMPLEX: interleaved (re,im) f32,
 * one output element = 8 bytes; all others: one f32 per bin.
 * MAG = |X| (norm(), beep/view.rs:155; equals the f32 sqrt-of-squares of
 * graph/utils.rs:53 up to rounding); POWER = |X|^2; DB = 20*log10|X|. */
enum {
    KOPEK_OUT_COMPLEX = 0, KOPEK_OUT_MAG = 1, KOPEK_OUT_POWER = 2, KOPEK_OUT_DB = 3,
    /* Fused band epilogues (n = 1024, HALF bins only): no spectrum is written; each frame yields
     * KOPEK_BAND_ROW = 10 f32: eight band means of scale*|X| followed by the peak pick of
     * examples/pong/main.rs:200-207 -- index of the strictly largest mean (as f32; 0 if none > 0)
     * and that mean.  BANDS_LOG: consecutive bands of 4,4,8,16,32,64,128,256 bins from bin 0
     * (examples/graph/utils.rs:85-116, examples/pong/utils.rs:62-94); BANDS_LOW: eight bands of
     * 3 bins from bin 20 (examples/pong/utils.rs:96-114).  scale: kopek_plan_set_band_scale
     * (10 000 in graph/utils.rs:56, 100 in pong/utils.rs:41; default 1). */
    KOPEK_OUT_BANDS_LOG = 4, KOPEK_OUT_BANDS_LOW = 5
};
#define KOPEK_BAND_ROW 10
/* HALF: bins 0..n/2 (n/2+1 per frame, real input makes the rest redundant);
 * FULL: all n bins, as the reference's callers see them (graph/utils.rs:50 iterates all): bins n/2+1 .. n-1 are the
 * conjugate mirror of a real frame, written by the same kernel beside the computed half.
 * Every combination of n, hop (any value >= 1), window, output and bins, and any 4-byte aligned sample pointer, runs on
 * the lane-group kernels; hops / bases that keep every frame 8-byte (n = 256: 16-byte) aligned take the vector loads. */
enum { KOPEK_BINS_HALF = 0, KOPEK_BINS_FULL = 1 };

/* Thread-local, NUL-terminated, valid until the next failing call on this thread. */
KOPEK_API const char *kopek_last_error(void);
KOPEK_API const char *kopek_version(void);
/* Number of visible CUDA devices (0 if none / driver missing). */
KOPEK_API int32_t kopek_device_count(void);

/* ---- 1. exact drop-in for kopek::fft::fft (src/fft.rs:6) --------------------
 * Host pointers, interleaved (re,im) f64 -- the memory layout of &[Complex<f64>].
 * Output length is kopek_next_pow2(n_in) (usize::next_power_of_two: 0 -> 1); the
 * input is zero-padded to it (src/fft.rs:31-36).  n_out must equal that length.
 * The device kernel performs the same radix-2 decimation-in-time butterflies
 * with the same twiddles, products and sums as src/fft.rs:23-27 without fused
 * multiply-add, so the result is bit-identical to the f64 recursion. */
KOPEK_API size_t kopek_next_pow2(size_t n);
KOPEK_API kopek_status kopek_fft_c2c_f64(const double *in, size_t n_in, double *out, size_t n_out);
/* `batch` independent transforms of n_in points each, laid out back to back
 * (input stride n_in, output stride kopek_next_pow2(n_in) complex values). */
/* Inverse of the above through the forward transform, x = conj(fft(conj(X))) / n (the reference has no inverse:
 * this is the one a caller of its fft() would write).  n must be a power of two; host pointers. */
KOPEK_API kopek_status kopek_ifft_c2c_f64(const double *in, size_t n, double *out);
KOPEK_API kopek_status kopek_fft_c2c_f64_batched(const double *in, size_t n_in, size_t batch, double *out);
/* Same, device-resident buffers, asynchronous on `cuda_stream` (a cudaStream_t; NULL = default). */
KOPEK_API kopek_status kopek_fft_c2c_f64_device(const double *d_in, size_t n_in, size_t batch, double *d_out,
                                                int device, void *cuda_stream);

/* kopek::fft::show (src/fft.rs:43-52): label line, then "{:.4}{:+.4}i" bins joined by ", ".
 * kopek_show_format writes at most cap-1 bytes + NUL into dst and returns the length needed
 * (snprintf convention; dst may be NULL when cap == 0).  kopek_show prints it to stdout. */
KOPEK_API size_t kopek_show_format(const char *label, const double *buf, size_t n, char *dst, size_t cap);
KOPEK_API kopek_status kopek_show(const char *label, const double *buf, size_t n);

/* ---- 2. throughput path: window -> real FFT -> epilogue, one kernel ----------
 * Frame f covers samples [f*hop, f*hop + n); frames = 1 + (n_samples - n)/hop
 * (0 if n_samples < n; no edge padding).  n is a power of two in [256, 4096].
 * Output row f starts at element f*out_pitch_elems (pitch >= bins per frame;
 * 0 = tightly packed).  Windows are periodic (DFT-even), built in f64, rounded
 * once to f32: hann w[i] = 0.5-0.5cos(2 pi i/n), hamming 0.54-0.46cos(2 pi i/n). */
typedef struct kopek_plan kopek_plan;
KOPEK_API kopek_status kopek_plan_create(kopek_plan **plan, uint32_t n, uint32_t hop, uint32_t window,
                                         uint32_t output, uint32_t bins, uint64_t out_pitch_elems, int device);
/* Opt-in epilogue for n = 1024 f32 spectra with 16-byte aligned rows (out_pitch_elems % 4 == 0, e.g. 516): each
 * frame leaves the SM as one 2048-byte bulk (TMA) store + one word instead of 32-bit stores.  Same speed on
 * local HBM; over NVLink (d_out mapped from a peer GPU, section 2d) it is what reaches link bandwidth.
 * Falls back to the direct stores when d_out is not 16-byte aligned. */
KOPEK_API kopek_status kopek_plan_set_bulk_store(kopek_plan *plan, int enable);
/* Display scale of the band epilogues (KOPEK_OUT_BANDS_*); takes effect at the next execute. */
KOPEK_API kopek_status kopek_plan_set_band_scale(kopek_plan *plan, float scale);
/* Replace the window by n caller-supplied f32 coefficients (host pointer). */
KOPEK_API kopek_status kopek_plan_set_window(kopek_plan *plan, const float *window_host);
KOPEK_API kopek_status kopek_plan_destroy(kopek_plan *plan);
KOPEK_API uint64_t kopek_plan_frames(const kopek_plan *plan, uint64_t n_samples);
KOPEK_API uint64_t kopek_plan_bins(const kopek_plan *plan);          /* per frame                */
KOPEK_API uint64_t kopek_plan_out_pitch(const kopek_plan *plan);     /* elements between rows    */
KOPEK_API uint64_t kopek_plan_out_elem_bytes(const kopek_plan *plan);/* 4, or 8 for COMPLEX      */
KOPEK_API uint32_t kopek_plan_last_launches(const kopek_plan *plan); /* kernels of the last exec */

/* Device-resident execute, asynchronous on cuda_stream.  d_samples: n_samples f32
 * on the plan's device; d_out: frames*pitch elements.  *n_frames_out (may be NULL)
 * receives the frame count. */
KOPEK_API kopek_status kopek_stft_exec(const kopek_plan *plan, const float *d_samples, uint64_t n_samples,
                                       void *d_out, uint64_t *n_frames_out, void *cuda_stream);
/* Host-buffer execute (the call a reference-side caller makes): copies chunks of
 * samples to the device, transforms, copies spectra back, overlapping the three on
 * internal streams; returns when h_out is complete.  Pinned buffers
 * (kopek_host_alloc) make the copies asynchronous; pageable memory also works. */
KOPEK_API kopek_status kopek_stft_exec_host(kopek_plan *plan, const float *h_samples, uint64_t n_samples,
                                            void *h_out, uint64_t *n_frames_out);
KOPEK_API kopek_status kopek_host_alloc(void **ptr, size_t bytes);   /* page-locked host memory */
KOPEK_API kopek_status kopek_host_free(void *ptr);

/* ---- 2b. live sliding window (examples/beep/view.rs:44,57-60,140-158) -----------------------
 * The beep view keeps the most recent 1024 samples in a VecDeque (new samples pushed at the back, old
 * ones popped at the front) and transforms the whole deque once per UI frame.  A kopek_stream is that
 * deque for the device (a host ring + a mapped pinned, aligned window which the kernel reads in place: one launch per
 * push, no DMA, the same kernel whatever the push sizes add up to): push() appends `count` new host samples (history
 * starts as zeros), then writes
 * the spectrum of the most recent n samples to h_out (kopek_stream_bins() elements, f32 or (re,im) f32).
 * input_scale multiplies every sample before the transform (1/32767 in beep/view.rs:143, 1 in pong). */
typedef struct kopek_stream kopek_stream;
KOPEK_API kopek_status kopek_stream_create(kopek_stream **stream, uint32_t n, uint32_t window, uint32_t output,
                                           float input_scale, int device);
KOPEK_API kopek_status kopek_stream_push(kopek_stream *stream, const float *h_samples, uint32_t count, void *h_out);
KOPEK_API uint64_t kopek_stream_bins(const kopek_stream *stream);
KOPEK_API kopek_status kopek_stream_destroy(kopek_stream *stream);

/* ---- 2c. on-device signal sources (src/oscillator.rs:34-91, src/noise.rs:15-21) -------------
 * Writes n_frames frames of n f32 samples (frame f at d_out + f*stride).  Each frame is one Oscillator
 * started at phase 0 (Oscillator::new) running at freq0 + freq_step*((first_frame + f) % freq_mod) Hz
 * (freq_mod 0: every frame at freq0), wave: 0 sine, 1 fake_sine, 2 sawtooth, 3 square(duty), 4 triangle,
 * 5 silence; the f32 phase-accumulator recurrence of the reference is executed verbatim (the phase
 * sequence is bit-identical, sinf within 2 ulp of glibc).  noise: 0 none, 1 rand_noise U[-1,1),
 * 2 white_noise N(0,1), added as value += noise_gain * noise like examples/beep/generator.rs:62-70; the
 * reference's generator is an unseeded ThreadRng, so only the distribution is reproduced (counter
 * based, a function of seed and the global sample index: shards generate the same signal). */
KOPEK_API kopek_status kopek_synth_frames(float *d_out, uint64_t n_frames, uint32_t n, uint64_t stride,
                                          float sample_rate, uint32_t wave, float duty, float freq0, float freq_step,
                                          uint32_t freq_mod, uint32_t noise, float noise_gain, uint64_t seed,
                                          uint64_t first_frame, int device, void *cuda_stream);

/* ---- 2c'. decoder-side callers on i16 PCM frames (kopek::decoder::decode returns Vec<[i16; 2]>, src/decoder.rs:1-31) ----
 * kopek_pcm16_to_f32: the marshalling of examples/graph/player.rs:56-59 -- frame[channel] as f64 / divisor (i16::MAX
 * there), narrowed once to f32 -- from n_frames interleaved frames of `channels` i16 into the f32 sample buffer the
 * STFT plans transform.  Bit-exact.  d_pcm 2-byte, d_out 4-byte aligned (16-byte aligned stereo takes the vector path).
 * kopek_detect_bpm: detect_bpm of src/decoder.rs:39-71 on device-resident stereo frames (16-byte aligned): one-second
 * windows of 44100 frames (hard coded in the reference), 43 blocks of 1024 frames each, a block is a beat when its
 * energy exceeds 5.5 x the second's mean block energy; *bpm_out = beats * (60 / duration) in u32 arithmetic.  Every
 * f32 sum is folded in the reference's order, so block energies, averages, beats and bpm are bit-identical to it.
 * Where the Rust original panics (less than one second: 60 / 0; sample_rate < 44100: slice index out of range)
 * this returns KOPEK_ERR_INVALID.  d_block_e (duration*43 f32) and d_avg_e (duration f32) are optional outputs.
 * Synchronous: returns after the count has been read back. */
KOPEK_API kopek_status kopek_pcm16_to_f32(const int16_t *d_pcm, uint64_t n_frames, uint32_t channels, uint32_t channel,
                                          double divisor, float *d_out, int device, void *cuda_stream);
/* kopek_pcm16_scale_f32: load_track_at_path of examples/graph/player.rs:81-102 -- the decoded Vec<[i16; 2]> becomes the
 * interleaved f32 track [volume_factor * frame[0] as f32, volume_factor * frame[1] as f32, ...] (0.00002 there): one
 * f32 product per value, bit-exact.  n_values = frames * channels; d_pcm 2-byte, d_out 4-byte aligned (16-byte
 * aligned buffers take the 128-bit path); asynchronous on cuda_stream. */
KOPEK_API kopek_status kopek_pcm16_scale_f32(const int16_t *d_pcm, uint64_t n_values, float factor, float *d_out, int device,
                                             void *cuda_stream);
/* kopek_f32_downmix: the capture callback of examples/pong/audio.rs:57-70 -- every interleaved f32 frame of `channels`
 * samples becomes ONE sample: the first two channels summed from 0.0 in order, divided by the channel count (what the
 * game then pops 1024 of and hands to fft(), examples/pong/main.rs:175-196).  Bit-exact; asynchronous on cuda_stream. */
KOPEK_API kopek_status kopek_f32_downmix(const float *d_frames, uint64_t n_frames, uint32_t channels, float *d_out, int device,
                                         void *cuda_stream);
KOPEK_API kopek_status kopek_detect_bpm(const int16_t *d_frames, uint64_t n_frames, float sample_rate, uint32_t *bpm_out,
                                        uint32_t *beats_out, float *d_block_e, float *d_avg_e, int device,
                                        void *cuda_stream);
/* the drop-in shape of detect_bpm(frames: Vec<[i16; 2]>, sample_rate: f32) -> u32: HOST frames; the duration * 44100
 * frames the algorithm reads are staged to the device, the rest as above */
KOPEK_API kopek_status kopek_detect_bpm_host(const int16_t *h_frames, uint64_t n_frames, float sample_rate,
                                             uint32_t *bpm_out, uint32_t *beats_out, int device);

/* kopek_stft_exec_pcm16: kopek_stft_exec on decoded audio -- n_frames interleaved STEREO i16 frames, sample =
 * frame[channel] as f64 / i16::MAX (examples/graph/player.rs:56-59); hop and n count frames.  For n = 1024 plans
 * (HALF bins, hop % 4 == 0, 16-byte aligned frames) the marshalling is fused into the kernel's load stage: one
 * pass from PCM to spectrum, results bit-identical to kopek_pcm16_to_f32 followed by kopek_stft_exec; other plans
 * take exactly that two-step route through a stream-ordered scratch buffer. */
KOPEK_API kopek_status kopek_stft_exec_pcm16(const kopek_plan *plan, const int16_t *d_frames, uint64_t n_frames,
                                             uint32_t channel, void *d_out, uint64_t *n_frames_out, void *cuda_stream);

/* ---- 2d. peer-direct spectrogram (multi-GPU, one process per GPU) -------------------------------
 * When the caller wants ONE spectrogram on one GPU, the other ranks do not send their spectra after
 * the fact: they run kopek_stft_exec with d_out pointing INTO the owner's buffer, mapped through CUDA
 * IPC, so the epilogue's stores travel over NVLink as they are produced (transform and transfer are
 * one kernel).  The owner allocates with kopek_device_alloc and exports a 64-byte handle; peers open it. */
#define KOPEK_IPC_HANDLE_BYTES 64
KOPEK_API kopek_status kopek_device_alloc(void **d_ptr, size_t bytes, int device);
KOPEK_API kopek_status kopek_device_free(void *d_ptr, int device);
/* Copy between any two of host / device / peer-mapped memory (cudaMemcpyDefault), synchronous. */
KOPEK_API kopek_status kopek_copy(void *dst, const void *src, size_t bytes, int device);
KOPEK_API kopek_status kopek_ipc_export(const void *d_ptr, unsigned char *handle_out /* 64 bytes */);
KOPEK_API kopek_status kopek_ipc_open(const unsigned char *handle /* 64 bytes */, int device, void **d_ptr);
KOPEK_API kopek_status kopek_ipc_close(void *d_ptr, int device);

/* ---- 3. large single transform (four-step), device-resident c2c f32 ----------
 * Interleaved (re,im) f32, n a power of two with 2^16 <= n <= 2^24, natural-order unnormalised
 * output, d_in != d_out, both 16-byte aligned.  Two HBM passes with TMA-staged tiles; asynchronous
 * on cuda_stream.  Replaces one fft() call (src/fft.rs:6) on a long buffer. */
KOPEK_API kopek_status kopek_fft_large_c2c_f32(const float *d_in, float *d_out, uint64_t n, int device,
                                               void *cuda_stream);

/* ---- 2e. one large transform over R GPUs (one process per GPU) -----------------------------------------
 * n = R * local_n points; rank r holds the cyclic slice x[j R + r].  Each rank runs kopek_fft_large_c2c_f32 on its
 * slice (F_r, local_n complex f32), the ranks exchange CUDA-IPC handles of those buffers (2d) and, after a barrier,
 * each calls kopek_fft_dist_combine: ONE kernel that reads F_r[k'] of every rank over NVLink for the k' of this
 * rank's block [rank * local_n / R, ...), applies W_n^{r k'} and the radix-R butterfly across the ranks, and writes
 * d_out[q][i] = X[(rank * local_n / R + i) + local_n * q]  (R rows of local_n / R complex).  d_spectra is a HOST
 * array of R device pointers (entry `rank` = this rank's own buffer, the others peer mappings); R = 2, 4 or 8. */
KOPEK_API kopek_status kopek_fft_dist_combine(const float *const *d_spectra, uint32_t ranks, uint32_t rank,
                                              uint64_t local_n, float *d_out, int device, void *cuda_stream);

/* ---- 4. multi-GPU in ONE process: frames sharded over R devices, optional gather to devices[0] -----------------
 * (what a Rust host links against -- no torch, no Python, one thread of the caller.)
 * The reference's fft is a pure function of one frame (src/fft.rs:6), so frames shard with NO data-path collective:
 * device r of R owns frames [floor(r F / R), floor((r + 1) F / R)) of the F = 1 + (n_samples - n) / hop frames of a
 * buffer and reads samples [f0 * hop, (f1 - 1) * hop + n) -- neighbours both read the (n - hop)-sample halo.
 * kopek_mgpu_shard reports those ranges; plan parameters are those of kopek_plan_create, one plan per device.
 *
 * kopek_mgpu_exec          device-resident shards: d_samples[r] / n_samples[r] / d_out[r] live on devices[r]; the R
 *                          launches are enqueued on the context's per-device streams and the call returns.
 * kopek_mgpu_exec_gathered same, but every device's kernel stores its rows straight INTO d_dst on devices[0] (rows in
 *                          rank order, pitch = kopek_mgpu_out_pitch) through the peer mapping: transform and transfer
 *                          are one kernel.  n = 1024 plans with f32 spectra and out_pitch_elems % 4 == 0 (e.g. 516)
 *                          use the 2048-byte bulk-store epilogue, which is what reaches NVLink bandwidth.
 * kopek_mgpu_gather        collects already computed shards (d_out[r], n_frames[r] rows each) into d_dst on devices[0]:
 *                          KOPEK_GATHER_COPY = copy-engine peer copies pushed by the source devices' streams,
 *                          KOPEK_GATHER_NCCL = one grouped ncclSend/ncclRecv (libnccl.so.2 is opened at run time;
 *                          KOPEK_ERR_UNSUPPORTED if it is not installed).
 * kopek_mgpu_sync          waits for everything enqueued so far; *elapsed_ms_max (may be NULL) = device time from the
 *                          first enqueue since the previous sync to the last, maximum over the devices (CUDA events).
 * kopek_mgpu_exec_host     ONE host buffer in, ONE host spectrogram out ([F][pitch]): the library splits the frame
 *                          ranges and runs each device's chunked H2D -> kernel -> D2H pipeline from its own host
 *                          thread; blocking.  Results are bit-identical to the one-device call (same kernels, same
 *                          frames).  Pinned buffers (kopek_host_alloc) make the copies asynchronous. */
typedef struct kopek_mgpu kopek_mgpu;
enum { KOPEK_GATHER_COPY = 0, KOPEK_GATHER_NCCL = 1 };
KOPEK_API kopek_status kopek_mgpu_create(kopek_mgpu **ctx, const int32_t *devices, uint32_t n_devices, uint32_t n,
                                         uint32_t hop, uint32_t window, uint32_t output, uint32_t bins,
                                         uint64_t out_pitch_elems);
KOPEK_API kopek_status kopek_mgpu_destroy(kopek_mgpu *ctx);
KOPEK_API uint32_t kopek_mgpu_device_count(const kopek_mgpu *ctx);
KOPEK_API uint64_t kopek_mgpu_bins(const kopek_mgpu *ctx);
KOPEK_API uint64_t kopek_mgpu_out_pitch(const kopek_mgpu *ctx);
KOPEK_API uint64_t kopek_mgpu_out_elem_bytes(const kopek_mgpu *ctx);
/* KOPEK_WINDOW_CUSTOM contexts: n coefficients (host pointer) for every device's plan */
KOPEK_API kopek_status kopek_mgpu_set_custom_window(kopek_mgpu *ctx, const float *window_host);
KOPEK_API kopek_status kopek_mgpu_shard(const kopek_mgpu *ctx, uint64_t n_samples, uint32_t rank,
                                        uint64_t *sample_begin, uint64_t *sample_count, uint64_t *frame_begin,
                                        uint64_t *frame_count);
KOPEK_API kopek_status kopek_mgpu_exec(kopek_mgpu *ctx, const float *const *d_samples, const uint64_t *n_samples,
                                       void *const *d_out, uint64_t *n_frames_out /* R entries, may be NULL */);
KOPEK_API kopek_status kopek_mgpu_exec_gathered(kopek_mgpu *ctx, const float *const *d_samples,
                                                const uint64_t *n_samples, void *d_dst, uint64_t *n_frames_total);
KOPEK_API kopek_status kopek_mgpu_gather(kopek_mgpu *ctx, void *const *d_out, const uint64_t *n_frames, void *d_dst,
                                         uint32_t mode);
KOPEK_API kopek_status kopek_mgpu_sync(kopek_mgpu *ctx, float *elapsed_ms_max);
KOPEK_API kopek_status kopek_mgpu_exec_host(kopek_mgpu *ctx, const float *h_samples, uint64_t n_samples, void *h_out,
                                            uint64_t *n_frames_out);

#ifdef __cplusplus
}
#endif
#endif /* KOPEK_CUDA_H */
