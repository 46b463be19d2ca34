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

// kopek::fft::stft_magnitude / stft_magnitude_multi of rust/kopek/src/fft.rs: many frames per call through the fused
// window -> real FFT -> |X| plan, on one device or sharded over several from this one process (kopek_mgpu_*).
inline std::vector<std::vector<float>> stft_magnitude(const std::vector<float> &samples, uint32_t n, uint32_t hop, bool hann) {
    kopek_plan *plan = nullptr;
    if (kopek_plan_create(&plan, n, hop, hann ? KOPEK_WINDOW_HANN : KOPEK_WINDOW_RECT, KOPEK_OUT_MAG, KOPEK_BINS_HALF, 0,
                          0) != KOPEK_OK)
        throw std::runtime_error(std::string("kopek::fft::stft_magnitude: ") + kopek_last_error());
    const size_t frames = kopek_plan_frames(plan, samples.size()), bins = kopek_plan_bins(plan);
    std::vector<float> flat(frames * bins);
    uint64_t got = 0;
    kopek_status st = kopek_stft_exec_host(plan, samples.data(), samples.size(), flat.data(), &got);
    kopek_plan_destroy(plan);
    if (st != KOPEK_OK) throw std::runtime_error(std::string("kopek::fft::stft_magnitude: ") + kopek_last_error());
    std::vector<std::vector<float>> rows(frames);
    for (size_t f = 0; f < frames; ++f) rows[f].assign(flat.begin() + f * bins, flat.begin() + (f + 1) * bins);
    return rows;
}

inline std::vector<std::vector<float>> stft_magnitude_multi(const std::vector<float> &samples, uint32_t n, uint32_t hop,
                                                            bool hann, const std::vector<int32_t> &devices) {
    kopek_mgpu *ctx = nullptr;
    if (kopek_mgpu_create(&ctx, devices.data(), (uint32_t)devices.size(), n, hop, hann ? KOPEK_WINDOW_HANN : KOPEK_WINDOW_RECT,
                          KOPEK_OUT_MAG, KOPEK_BINS_HALF, 0) != KOPEK_OK)
        throw std::runtime_error(std::string("kopek::fft::stft_magnitude_multi: ") + kopek_last_error());
    const size_t bins = kopek_mgpu_bins(ctx);
    const size_t frames = samples.size() < n ? 0 : 1 + (samples.size() - n) / hop;
    std::vector<float> flat(frames * bins);
    uint64_t got = 0;
    kopek_status st = kopek_mgpu_exec_host(ctx, samples.data(), samples.size(), flat.data(), &got);
    std::string msg = st != KOPEK_OK ? kopek_last_error() : "";
    kopek_mgpu_destroy(ctx);
    if (st != KOPEK_OK) throw std::runtime_error("kopek::fft::stft_magnitude_multi: " + msg);
    std::vector<std::vector<float>> rows(frames);
    for (size_t f = 0; f < frames; ++f) rows[f].assign(flat.begin() + f * bins, flat.begin() + (f + 1) * bins);
    return rows;
}

}  // namespace fft
}  // namespace kopek
