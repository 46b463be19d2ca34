//! Drop-in for /root/reference/src/fft.rs: identical signatures, the transform runs on the GPU
//! through libkopek_cuda.so.  `fft` stays infallible like the original: a CUDA failure panics
//! with the library's error text (there is deliberately no CPU fallback).
use kopek_cuda_sys as sys;
use num::complex::Complex;
use std::ffi::{CStr, CString};

fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::kopek_last_error()) }.to_string_lossy().into_owned()
}

/// Forward FFT, zero-padded to the next power of two (len 0 -> 1), natural order, unnormalised.
/// Bit-identical to the reference's recursive radix-2 implementation (src/fft.rs:6-41).
pub fn fft(input: &[Complex<f64>]) -> Vec<Complex<f64>> {
    let n = input.len().next_power_of_two();
    let mut out = vec![Complex { re: 0.0, im: 0.0 }; n];
    // Complex<f64> is #[repr(C)] { re, im }: a slice of it is n interleaved (re, im) doubles.
    let st = unsafe {
        sys::kopek_fft_c2c_f64(input.as_ptr() as *const f64, input.len(), out.as_mut_ptr() as *mut f64, n)
    };
    if st != sys::KOPEK_OK {
        panic!("kopek::fft::fft: {}", last_error());
    }
    out
}

/// Prints the label and the bins as "{:.4}{:+.4}i" joined by ", " (src/fft.rs:43-52).
pub fn show(label: &str, buf: &[Complex<f64>]) {
    let c = CString::new(label).expect("label contains NUL");
    let st = unsafe { sys::kopek_show(c.as_ptr(), buf.as_ptr() as *const f64, buf.len()) };
    if st != sys::KOPEK_OK {
        panic!("kopek::fft::show: {}", last_error());
    }
}

/// Batched analysis of many frames at once (what the three callers of `fft` -- examples/beep/view.rs:140-158,
/// examples/graph/player.rs:56-61, examples/pong/main.rs:192-196 -- become when more than one frame is pending):
/// window -> real FFT -> |X| in one fused kernel.  `samples` are f32 as the callers hold them.
pub fn stft_magnitude(samples: &[f32], n: u32, hop: u32, hann: bool) -> Vec<Vec<f32>> {
    let mut plan: *mut sys::kopek_plan = std::ptr::null_mut();
    let window = if hann { sys::KOPEK_WINDOW_HANN } else { sys::KOPEK_WINDOW_RECT };
    let st = unsafe { sys::kopek_plan_create(&mut plan, n, hop, window, sys::KOPEK_OUT_MAG, sys::KOPEK_BINS_HALF, 0, 0) };
    if st != sys::KOPEK_OK {
        panic!("kopek::fft::stft_magnitude: {}", last_error());
    }
    let frames = unsafe { sys::kopek_plan_frames(plan, samples.len() as u64) } as usize;
    let bins = unsafe { sys::kopek_plan_bins(plan) } as usize;
    let mut flat = vec![0f32; frames * bins];
    let mut got = 0u64;
    let st = unsafe {
        sys::kopek_stft_exec_host(plan, samples.as_ptr(), samples.len() as u64, flat.as_mut_ptr() as *mut _, &mut got)
    };
    unsafe { sys::kopek_plan_destroy(plan) };
    if st != sys::KOPEK_OK {
        panic!("kopek::fft::stft_magnitude: {}", last_error());
    }
    flat.chunks(bins).map(|r| r.to_vec()).collect()
}
