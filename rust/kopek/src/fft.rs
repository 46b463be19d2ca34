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

/// A fused window -> real FFT -> epilogue plan (kopek_plan_*): device tables and the chunk slots of the host-buffer
/// path live as long as the value, so a caller that analyses a stream of buffers builds it once.
pub struct Plan {
    raw: *mut sys::kopek_plan,
    bins: usize,
}
// the C library serialises host-buffer executes of one plan behind its own mutex
unsafe impl Send for Plan {}

impl Plan {
    pub fn new(n: u32, hop: u32, window: u32, output: u32, device: i32) -> Plan {
        let mut raw: *mut sys::kopek_plan = std::ptr::null_mut();
        let st = unsafe { sys::kopek_plan_create(&mut raw, n, hop, window, output, sys::KOPEK_BINS_HALF, 0, device) };
        if st != sys::KOPEK_OK {
            panic!("kopek::fft::Plan::new: {}", last_error());
        }
        let bins = unsafe { sys::kopek_plan_bins(raw) } as usize;
        Plan { raw, bins }
    }

    pub fn bins(&self) -> usize {
        self.bins
    }

    /// Host samples in, `frames x bins` f32 values out (row-major); copies happen inside the call.
    pub fn exec(&mut self, samples: &[f32]) -> Vec<f32> {
        let frames = unsafe { sys::kopek_plan_frames(self.raw, samples.len() as u64) } as usize;
        let mut flat = vec![0f32; frames * self.bins];
        let mut got = 0u64;
        let st = unsafe {
            sys::kopek_stft_exec_host(self.raw, samples.as_ptr(), samples.len() as u64, flat.as_mut_ptr() as *mut _, &mut got)
        };
        if st != sys::KOPEK_OK {
            panic!("kopek::fft::Plan::exec: {}", last_error());
        }
        flat
    }
}

impl Drop for Plan {
    fn drop(&mut self) {
        unsafe { sys::kopek_plan_destroy(self.raw) };
    }
}

thread_local! {
    // the few most recently used plans of this thread, keyed by (n, hop, window): a UI loop that calls
    // stft_magnitude once per rendered frame pays for cudaMalloc and the table uploads once, not per call
    static PLANS: std::cell::RefCell<Vec<((u32, u32, u32), Plan)>> = std::cell::RefCell::new(Vec::new());
}
const PLAN_CACHE: usize = 4;

/// Batched analysis of many frames at once (what the three callers of `fft` -- examples/beep/view.rs:140-158,
/// examples/graph/player.rs:56-61, examples/pong/main.rs:192-196 -- become when more than one frame is pending):
/// window -> real FFT -> |X| in one fused kernel.  `samples` are f32 as the callers hold them.
pub fn stft_magnitude(samples: &[f32], n: u32, hop: u32, hann: bool) -> Vec<Vec<f32>> {
    let window = if hann { sys::KOPEK_WINDOW_HANN } else { sys::KOPEK_WINDOW_RECT };
    PLANS.with(|cell| {
        let mut cache = cell.borrow_mut();
        let key = (n, hop, window);
        let at = match cache.iter().position(|(k, _)| *k == key) {
            Some(i) => i,
            None => {
                if cache.len() >= PLAN_CACHE {
                    cache.remove(0); // least recently used sits in front
                }
                cache.push((key, Plan::new(n, hop, window, sys::KOPEK_OUT_MAG, 0)));
                cache.len() - 1
            }
        };
        let entry = cache.remove(at);
        cache.push(entry); // most recently used at the back
        let plan = &mut cache.last_mut().unwrap().1;
        let bins = plan.bins();
        plan.exec(samples).chunks(bins).map(|r| r.to_vec()).collect()
    })
}

/// The same analysis spread over several GPUs of this machine from ONE process (kopek_mgpu_*): the library splits the
/// frame range, every device transforms its contiguous share (frames are independent, src/fft.rs:6 is a pure function)
/// and the rows land at their final place in the result, bit-identical to the single-device call.
pub fn stft_magnitude_multi(samples: &[f32], n: u32, hop: u32, hann: bool, devices: &[i32]) -> Vec<Vec<f32>> {
    let window = if hann { sys::KOPEK_WINDOW_HANN } else { sys::KOPEK_WINDOW_RECT };
    let mut ctx: *mut sys::kopek_mgpu = std::ptr::null_mut();
    let st = unsafe {
        sys::kopek_mgpu_create(&mut ctx, devices.as_ptr(), devices.len() as u32, n, hop, window, sys::KOPEK_OUT_MAG,
                               sys::KOPEK_BINS_HALF, 0)
    };
    if st != sys::KOPEK_OK {
        panic!("kopek::fft::stft_magnitude_multi: {}", last_error());
    }
    let bins = unsafe { sys::kopek_mgpu_bins(ctx) } as usize;
    let frames = if samples.len() < n as usize { 0 } else { 1 + (samples.len() - n as usize) / hop as usize };
    let mut flat = vec![0f32; frames * bins];
    let mut got = 0u64;
    let st = unsafe {
        sys::kopek_mgpu_exec_host(ctx, samples.as_ptr(), samples.len() as u64, flat.as_mut_ptr() as *mut _, &mut got)
    };
    let msg = if st != sys::KOPEK_OK { Some(last_error()) } else { None };
    unsafe { sys::kopek_mgpu_destroy(ctx) };
    if let Some(m) = msg {
        panic!("kopek::fft::stft_magnitude_multi: {}", m);
    }
    flat.chunks(bins).map(|r| r.to_vec()).collect()
}

/// The sliding window of the live examples (examples/beep/view.rs:44,57-60,140-158): the `VecDeque` of the most recent
/// `n` samples lives in the library (kopek_stream_*); `push` appends the samples that arrived since the last UI frame
/// and returns the magnitude spectrum (n/2 + 1 bins) of the window that ends at the newest one.  `input_scale` is the
/// `/ 32767.0` of beep/view.rs:143 (1.0 in pong).
pub struct LiveSpectrum {
    raw: *mut sys::kopek_stream,
    bins: usize,
}
unsafe impl Send for LiveSpectrum {}

impl LiveSpectrum {
    pub fn new(n: u32, hann: bool, input_scale: f32) -> LiveSpectrum {
        let mut raw: *mut sys::kopek_stream = std::ptr::null_mut();
        let window = if hann { sys::KOPEK_WINDOW_HANN } else { sys::KOPEK_WINDOW_RECT };
        let st = unsafe { sys::kopek_stream_create(&mut raw, n, window, sys::KOPEK_OUT_MAG, input_scale, 0) };
        if st != sys::KOPEK_OK {
            panic!("kopek::fft::LiveSpectrum::new: {}", last_error());
        }
        let bins = unsafe { sys::kopek_stream_bins(raw) } as usize;
        LiveSpectrum { raw, bins }
    }

    pub fn push(&mut self, new_samples: &[f32]) -> Vec<f32> {
        let mut out = vec![0f32; self.bins];
        let st = unsafe {
            sys::kopek_stream_push(self.raw, new_samples.as_ptr(), new_samples.len() as u32, out.as_mut_ptr() as *mut _)
        };
        if st != sys::KOPEK_OK {
            panic!("kopek::fft::LiveSpectrum::push: {}", last_error());
        }
        out
    }
}

impl Drop for LiveSpectrum {
    fn drop(&mut self) {
        unsafe { sys::kopek_stream_destroy(self.raw) };
    }
}
