// Host-side check of the register butterflies (Dft<2|4|8|16|64>::run in kopek_b200/csrc/fft_batched.cuh) against a
// direct f64 DFT.  The templates are __host__ __device__, so the very code the kernels unroll runs here on the CPU
// (nvcc host compilation, no GPU needed).  Exit code 0 = every radix within 2e-7 of the f64 result.
#include <cmath>
#include <complex>
#include <cstdio>

#include "../../kopek_b200/csrc/fft_batched.cuh"

template <int R> static double check() {
    using namespace kopek;
    float2 v[R];
    std::complex<double> x[R];
    for (int i = 0; i < R; ++i) {
        v[i] = make_float2(sinf(1.3f * i + 0.1f * R) + 0.01f * i, cosf(0.7f * i) - 0.02f * i);
        x[i] = {v[i].x, v[i].y};
    }
    Dft<R>::run(v);
    double err = 0, mx = 0;
    for (int k = 0; k < R; ++k) {
        std::complex<double> s = 0;
        for (int n = 0; n < R; ++n) s += x[n] * std::polar(1.0, -2.0 * M_PI * (double)(n * k % R) / R);
        err = fmax(err, std::abs(s - std::complex<double>(v[k].x, v[k].y)));
        mx = fmax(mx, std::abs(s));
    }
    printf("radix %2d: max error %.3g of %.3g (%.2e relative)\n", R, err, mx, err / mx);
    return err / mx;
}

int main() {
    double worst = 0;
    worst = fmax(worst, check<2>());
    worst = fmax(worst, check<4>());
    worst = fmax(worst, check<8>());
    worst = fmax(worst, check<16>());
    worst = fmax(worst, check<64>());
    return worst < 2e-7 ? 0 : 1;
}
