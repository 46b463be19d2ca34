// C++ harness of the drop-in boundary: the call sequences rust/kopek/src/fft.rs makes, through the
// identical C ABI (the Rust shim cannot be compiled in this image).  Links libkopek_cuda.so and the
// oracle (test infrastructure) and compares bit for bit.  Built and run by tests/test_gpu_dropin.py.
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../kopek_b200/host/kopek_fft.hpp"

extern "C" size_t oracle_fft_f64(const double *in, size_t n_in, double *out);
extern "C" void oracle_osc_sine(float sample_rate, float frequency, float *phase_io, float *out, size_t count);

int main() {
    int bad = 0;
    // BASELINE.json configs[0]: Oscillator::new(44100), 440 Hz, 1024 x sine(), widened like pong/main.rs:192-195
    std::vector<float> s(1024);
    float phase = 0.0f;
    oracle_osc_sine(44100.0f, 440.0f, &phase, s.data(), s.size());
    std::vector<std::complex<double>> in(1024);
    for (size_t i = 0; i < s.size(); ++i) in[i] = std::complex<double>((double)s[i], 0.0);
    std::vector<std::complex<double>> out = kopek::fft::fft(in);
    std::vector<std::complex<double>> ref(1024);
    oracle_fft_f64(reinterpret_cast<const double *>(in.data()), in.size(), reinterpret_cast<double *>(ref.data()));
    if (out.size() != 1024 || memcmp(out.data(), ref.data(), 1024 * sizeof(ref[0])) != 0) { printf("sine440 mismatch\n"); ++bad; }
    size_t peak = 0;
    for (size_t k = 1; k <= 512; ++k) if (std::abs(out[k]) > std::abs(out[peak])) peak = k;
    printf("peak bin %zu |X| %.4f\n", peak, std::abs(out[peak]));
    if (peak != 10) ++bad;
    // ragged lengths incl. 0 and 1 (src/fft.rs:31-36)
    for (size_t n_in : {size_t(0), size_t(1), size_t(3), size_t(1000), size_t(1025)}) {
        std::vector<std::complex<double>> x(n_in);
        for (size_t i = 0; i < n_in; ++i) x[i] = std::complex<double>(sin(0.37 * i), cos(1.3 * i));
        std::vector<std::complex<double>> y = kopek::fft::fft(x);
        std::vector<std::complex<double>> r(kopek_next_pow2(n_in));
        oracle_fft_f64(reinterpret_cast<const double *>(x.data()), n_in, reinterpret_cast<double *>(r.data()));
        if (y.size() != r.size() || memcmp(y.data(), r.data(), r.size() * sizeof(r[0])) != 0) { printf("n_in=%zu mismatch\n", n_in); ++bad; }
    }
    // many frames at once, one device and every visible device from this one process (kopek_mgpu_*): identical bits
    {
        std::vector<float> sig(48000);
        for (size_t i = 0; i < sig.size(); ++i) sig[i] = (float)sin(2.0 * M_PI * 440.0 * (double)i / 48000.0) + 0.25f * (float)cos(0.01 * i * i);
        std::vector<std::vector<float>> one = kopek::fft::stft_magnitude(sig, 1024, 256, true);
        std::vector<int32_t> devs;
        for (int d = 0; d < kopek_device_count() && d < 8; ++d) devs.push_back(d);
        std::vector<std::vector<float>> many = kopek::fft::stft_magnitude_multi(sig, 1024, 256, true, devs);
        bool same = one.size() == many.size() && one.size() == 1 + (sig.size() - 1024) / 256;
        for (size_t f = 0; same && f < one.size(); ++f)
            same = one[f].size() == 513 && memcmp(one[f].data(), many[f].data(), 513 * sizeof(float)) == 0;
        printf("stft_magnitude_multi over %zu device(s): %s\n", devs.size(), same ? "identical" : "MISMATCH");
        if (!same) ++bad;
    }
    kopek::fft::show("spectrum", std::vector<std::complex<double>>(out.begin(), out.begin() + 3));
    printf(bad ? "DROPIN_FAIL\n" : "DROPIN_OK\n");
    return bad;
}
