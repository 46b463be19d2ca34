pub mod decoder;
pub mod fft;
