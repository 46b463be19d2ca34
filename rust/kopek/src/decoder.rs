//! Drop-in for the energy pass of /root/reference/src/decoder.rs: `detect_bpm` keeps its signature and runs on the
//! GPU through libkopek_cuda.so (bit-identical beats and bpm: every f32 sum is folded in the original's order).
//! `decode` and `duration_in_seconds` stay what they are in the reference (file I/O through `audrey`, one division).
use kopek_cuda_sys as sys;
use std::ffi::CStr;

fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::kopek_last_error()) }.to_string_lossy().into_owned()
}

/// src/decoder.rs:39-71.  The original panics for less than one second of audio (`60 / 0`) and when
/// `sample_rate < 44100` (slice index out of range); so does this one, with the library's message.
pub fn detect_bpm(frames: Vec<[i16; 2]>, sample_rate: f32) -> u32 {
    let (mut bpm, mut beats) = (0u32, 0u32);
    // [i16; 2] is two adjacent i16: a Vec of it is n interleaved stereo frames
    let st = unsafe {
        sys::kopek_detect_bpm_host(frames.as_ptr() as *const i16, frames.len() as u64, sample_rate, &mut bpm, &mut beats, 0)
    };
    if st != sys::KOPEK_OK {
        panic!("kopek::decoder::detect_bpm: {}", last_error());
    }
    bpm
}
