// kopek_fft.hpp -- C++ host mirror of the reference's `kopek::fft` module over the C ABI.
//
// The reference is a Rust crate (/root/reference/src/fft.rs); this image has no Rust
// toolchain, so the host side above the C ABI is mirrored in C++ with the same names,
// argument meaning and error behaviour:
//
//   kopek::fft::fft(input)        <- pub fn fft(input: &[Complex<f64>]) -> Vec<Complex<f64>>   src/fft.rs:6
//   kopek::fft::show(label, buf)  <- pub fn show(label: &str, buf: &[Complex<f64>])            src/fft.rs:43
//
// `fft` is infallible in the reference (no Result, no documented panic).  There is no CPU
// fallback here, so a CUDA failure throws std::runtime_error carrying kopek_last_error()
// -- the analogue of the panic!() in rust/kopek/src/fft.rs.
// std::complex<double> is layout-compatible with num_complex::Complex<f64> (#[repr(C)] {re, im}).
#pragma once
#include <complex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/kopek_cuda.h"

namespace kopek {
namespace fft {

inline std::vector<std::complex<double>> fft(const std::vector<std::complex<double>> &input) {
    const size_t n = kopek_next_pow2(input.size());          // src/fft.rs:31-32 (len 0 -> 1)
    std::vector<std::complex<double>> out(n);
    kopek_status st = kopek_fft_c2c_f64(reinterpret_cast<const double *>(input.data()), input.size(),
                                        reinterpret_cast<double *>(out.data()), n);
    if (st != KOPEK_OK) throw std::runtime_error(std::string("kopek::fft::fft: ") + kopek_last_error());
    return out;
}

inline void show(const std::string &label, const std::vector<std::complex<double>> &buf) {
    kopek_status st = kopek_show(label.c_str(), reinterpret_cast<const double *>(buf.data()), buf.size());
    if (st != KOPEK_OK) throw std::runtime_error(std::string("kopek::fft::show: ") + kopek_last_error());
}

}  // namespace fft
}  // namespace kopek
