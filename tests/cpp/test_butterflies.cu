// Host-side check of the radix-2/4/8/16/64 register butterflies (K1 family, K3 pass 2) against a
// direct f64 DFT.  Built and run by tests/test_host_logic.py (no GPU needed).
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../../kopek_b200/csrc/fft_batched.cuh"

template <int R> static double check() {
    float2 v[R];
    double xr[R], xi[R];
    srand(7 + R);
    for (int i = 0; i < R; ++i) {
        xr[i] = rand() / (double)RAND_MAX - 0.5;
        xi[i] = rand() / (double)RAND_MAX - 0.5;
        v[i] = make_float2((float)xr[i], (float)xi[i]);
        xr[i] = v[i].x;
        xi[i] = v[i].y;
    }
    kopek::Dft<R>::run(v);
    double worst = 0;
    for (int k = 0; k < R; ++k) {
        double sr = 0, si = 0;
        for (int t = 0; t < R; ++t) {
            double th = -2.0 * M_PI * ((t * k) % R) / R;
            sr += xr[t] * cos(th) - xi[t] * sin(th);
            si += xr[t] * sin(th) + xi[t] * cos(th);
        }
        worst = fmax(worst, fmax(fabs(sr - v[k].x), fabs(si - v[k].y)));
    }
    return worst;
}

int main() {
    double e2 = check<2>(), e4 = check<4>(), e8 = check<8>(), e16 = check<16>(), e64 = check<64>();
    printf("dft2 %.3e dft4 %.3e dft8 %.3e dft16 %.3e dft64 %.3e\n", e2, e4, e8, e16, e64);
    return (e2 < 1e-6 && e4 < 2e-6 && e8 < 2e-6 && e16 < 4e-6 && e64 < 4e-6) ? 0 : 1;
}
