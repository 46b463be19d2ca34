//! Raw FFI to libkopek_cuda.so -- one declaration per entry point of include/kopek_cuda.h.
//! NOTE: written against the header; this repository's build image has no rustc/cargo, so
//! these sources are not compiled here.  tests/cpp/test_dropin.cpp exercises the identical
//! call sequences through the same C ABI.
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_int, c_void};

pub type kopek_status = i32;
pub const KOPEK_OK: kopek_status = 0;
pub const KOPEK_ERR_INVALID: kopek_status = -1;
pub const KOPEK_ERR_CUDA: kopek_status = -2;
pub const KOPEK_ERR_NOMEM: kopek_status = -3;
pub const KOPEK_ERR_UNSUPPORTED: kopek_status = -4;

pub const KOPEK_WINDOW_RECT: u32 = 0;
pub const KOPEK_WINDOW_HANN: u32 = 1;
pub const KOPEK_WINDOW_HAMMING: u32 = 2;
pub const KOPEK_WINDOW_CUSTOM: u32 = 3;
pub const KOPEK_OUT_COMPLEX: u32 = 0;
pub const KOPEK_OUT_MAG: u32 = 1;
pub const KOPEK_OUT_POWER: u32 = 2;
pub const KOPEK_OUT_DB: u32 = 3;
pub const KOPEK_OUT_BANDS_LOG: u32 = 4;
pub const KOPEK_OUT_BANDS_LOW: u32 = 5;
pub const KOPEK_BAND_ROW: usize = 10;
pub const KOPEK_BINS_HALF: u32 = 0;
pub const KOPEK_BINS_FULL: u32 = 1;

#[repr(C)]
pub struct kopek_plan {
    _private: [u8; 0],
}
#[repr(C)]
pub struct kopek_stream {
    _private: [u8; 0],
}
#[repr(C)]
pub struct kopek_mgpu {
    _private: [u8; 0],
}
pub const KOPEK_GATHER_COPY: u32 = 0;
pub const KOPEK_GATHER_NCCL: u32 = 1;

extern "C" {
    pub fn kopek_last_error() -> *const c_char;
    pub fn kopek_version() -> *const c_char;
    pub fn kopek_device_count() -> i32;

    pub fn kopek_next_pow2(n: usize) -> usize;
    /// `in_`/`out`: interleaved (re, im) f64 == `*const Complex<f64>` (num-complex is #[repr(C)]).
    pub fn kopek_fft_c2c_f64(in_: *const f64, n_in: usize, out: *mut f64, n_out: usize) -> kopek_status;
    pub fn kopek_ifft_c2c_f64(in_: *const f64, n: usize, out: *mut f64) -> kopek_status;
    pub fn kopek_fft_c2c_f64_batched(in_: *const f64, n_in: usize, batch: usize, out: *mut f64) -> kopek_status;
    pub fn kopek_fft_c2c_f64_device(d_in: *const f64, n_in: usize, batch: usize, d_out: *mut f64, device: c_int,
                                    cuda_stream: *mut c_void) -> kopek_status;
    pub fn kopek_show_format(label: *const c_char, buf: *const f64, n: usize, dst: *mut c_char, cap: usize) -> usize;
    pub fn kopek_show(label: *const c_char, buf: *const f64, n: usize) -> kopek_status;

    pub fn kopek_plan_create(plan: *mut *mut kopek_plan, n: u32, hop: u32, window: u32, output: u32, bins: u32,
                             out_pitch_elems: u64, device: c_int) -> kopek_status;
    pub fn kopek_plan_set_window(plan: *mut kopek_plan, window_host: *const f32) -> kopek_status;
    pub fn kopek_plan_set_band_scale(plan: *mut kopek_plan, scale: f32) -> kopek_status;
    pub fn kopek_plan_set_bulk_store(plan: *mut kopek_plan, enable: c_int) -> kopek_status;
    pub fn kopek_plan_destroy(plan: *mut kopek_plan) -> kopek_status;
    pub fn kopek_plan_frames(plan: *const kopek_plan, n_samples: u64) -> u64;
    pub fn kopek_plan_bins(plan: *const kopek_plan) -> u64;
    pub fn kopek_plan_out_pitch(plan: *const kopek_plan) -> u64;
    pub fn kopek_plan_out_elem_bytes(plan: *const kopek_plan) -> u64;
    pub fn kopek_plan_last_launches(plan: *const kopek_plan) -> u32;
    pub fn kopek_stft_exec(plan: *const kopek_plan, d_samples: *const f32, n_samples: u64, d_out: *mut c_void,
                           n_frames_out: *mut u64, cuda_stream: *mut c_void) -> kopek_status;
    pub fn kopek_stft_exec_host(plan: *mut kopek_plan, h_samples: *const f32, n_samples: u64, h_out: *mut c_void,
                                n_frames_out: *mut u64) -> kopek_status;
    pub fn kopek_stream_create(stream: *mut *mut kopek_stream, n: u32, window: u32, output: u32, input_scale: f32,
                               device: c_int) -> kopek_status;
    pub fn kopek_stream_push(stream: *mut kopek_stream, h_samples: *const f32, count: u32, h_out: *mut c_void)
                             -> kopek_status;
    pub fn kopek_stream_bins(stream: *const kopek_stream) -> u64;
    pub fn kopek_stream_destroy(stream: *mut kopek_stream) -> kopek_status;
    pub fn kopek_synth_frames(d_out: *mut f32, n_frames: u64, n: u32, stride: u64, sample_rate: f32, wave: u32,
                              duty: f32, freq0: f32, freq_step: f32, freq_mod: u32, noise: u32, noise_gain: f32,
                              seed: u64, first_frame: u64, device: c_int, cuda_stream: *mut c_void) -> kopek_status;
    pub fn kopek_stft_exec_pcm16(plan: *const kopek_plan, d_frames: *const i16, n_frames: u64, channel: u32,
                                 d_out: *mut c_void, n_frames_out: *mut u64, cuda_stream: *mut c_void) -> kopek_status;
    pub fn kopek_pcm16_to_f32(d_pcm: *const i16, n_frames: u64, channels: u32, channel: u32, divisor: f64,
                              d_out: *mut f32, device: c_int, cuda_stream: *mut c_void) -> kopek_status;
    pub fn kopek_pcm16_scale_f32(d_pcm: *const i16, n_values: u64, factor: f32, d_out: *mut f32, device: c_int,
                                 cuda_stream: *mut c_void) -> kopek_status;
    pub fn kopek_f32_downmix(d_frames: *const f32, n_frames: u64, channels: u32, d_out: *mut f32, device: c_int,
                             cuda_stream: *mut c_void) -> kopek_status;
    pub fn kopek_detect_bpm(d_frames: *const i16, n_frames: u64, sample_rate: f32, bpm_out: *mut u32,
                            beats_out: *mut u32, d_block_e: *mut f32, d_avg_e: *mut f32, device: c_int,
                            cuda_stream: *mut c_void) -> kopek_status;
    pub fn kopek_detect_bpm_host(h_frames: *const i16, n_frames: u64, sample_rate: f32, bpm_out: *mut u32,
                                 beats_out: *mut u32, device: c_int) -> kopek_status;
    pub fn kopek_device_alloc(d_ptr: *mut *mut c_void, bytes: usize, device: c_int) -> kopek_status;
    pub fn kopek_device_free(d_ptr: *mut c_void, device: c_int) -> kopek_status;
    pub fn kopek_copy(dst: *mut c_void, src: *const c_void, bytes: usize, device: c_int) -> kopek_status;
    pub fn kopek_ipc_export(d_ptr: *const c_void, handle_out: *mut u8) -> kopek_status;
    pub fn kopek_ipc_open(handle: *const u8, device: c_int, d_ptr: *mut *mut c_void) -> kopek_status;
    pub fn kopek_ipc_close(d_ptr: *mut c_void, device: c_int) -> kopek_status;
    pub fn kopek_fft_dist_combine(d_spectra: *const *const f32, ranks: u32, rank: u32, local_n: u64, d_out: *mut f32,
                                  device: c_int, cuda_stream: *mut c_void) -> kopek_status;
    pub fn kopek_host_alloc(ptr: *mut *mut c_void, bytes: usize) -> kopek_status;
    pub fn kopek_host_free(ptr: *mut c_void) -> kopek_status;

    pub fn kopek_fft_large_c2c_f32(d_in: *const f32, d_out: *mut f32, n: u64, device: c_int,
                                   cuda_stream: *mut c_void) -> kopek_status;

    // ---- multi-GPU in one process (include/kopek_cuda.h section 4)
    pub fn kopek_mgpu_create(ctx: *mut *mut kopek_mgpu, devices: *const i32, n_devices: u32, n: u32, hop: u32,
                             window: u32, output: u32, bins: u32, out_pitch_elems: u64) -> kopek_status;
    pub fn kopek_mgpu_destroy(ctx: *mut kopek_mgpu) -> kopek_status;
    pub fn kopek_mgpu_device_count(ctx: *const kopek_mgpu) -> u32;
    pub fn kopek_mgpu_bins(ctx: *const kopek_mgpu) -> u64;
    pub fn kopek_mgpu_out_pitch(ctx: *const kopek_mgpu) -> u64;
    pub fn kopek_mgpu_out_elem_bytes(ctx: *const kopek_mgpu) -> u64;
    pub fn kopek_mgpu_set_custom_window(ctx: *mut kopek_mgpu, window_host: *const f32) -> kopek_status;
    pub fn kopek_mgpu_shard(ctx: *const kopek_mgpu, n_samples: u64, rank: u32, sample_begin: *mut u64,
                            sample_count: *mut u64, frame_begin: *mut u64, frame_count: *mut u64) -> kopek_status;
    pub fn kopek_mgpu_exec(ctx: *mut kopek_mgpu, d_samples: *const *const f32, n_samples: *const u64,
                           d_out: *const *mut c_void, n_frames_out: *mut u64) -> kopek_status;
    pub fn kopek_mgpu_exec_gathered(ctx: *mut kopek_mgpu, d_samples: *const *const f32, n_samples: *const u64,
                                    d_dst: *mut c_void, n_frames_total: *mut u64) -> kopek_status;
    pub fn kopek_mgpu_gather(ctx: *mut kopek_mgpu, d_out: *const *mut c_void, n_frames: *const u64, d_dst: *mut c_void,
                             mode: u32) -> kopek_status;
    pub fn kopek_mgpu_sync(ctx: *mut kopek_mgpu, elapsed_ms_max: *mut f32) -> kopek_status;
    pub fn kopek_mgpu_exec_host(ctx: *mut kopek_mgpu, h_samples: *const f32, n_samples: u64, h_out: *mut c_void,
                                n_frames_out: *mut u64) -> kopek_status;
}
