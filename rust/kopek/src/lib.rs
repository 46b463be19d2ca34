pub mod fft;
