"""configs[0]: host-to-host latency of one fft() drop-in call (K2) at the sizes the reference's callers use,
called through the C ABI directly (ctypes overhead included) and through the Python mirror."""
import ctypes
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from kopek_b200 import _lib, fft  # noqa: E402

L = _lib.load()
rng = np.random.default_rng(0)
for n in (256, 1000, 1024, 4096, 8192, 65536):
    x = rng.standard_normal(n).astype(np.complex128)
    npad = 1 << max(0, (n - 1).bit_length())
    out = np.empty(npad, np.complex128)
    pin, pout = x.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p)
    for _ in range(20):
        assert L.kopek_fft_c2c_f64(pin, n, pout, npad) == 0
    lat = []
    for _ in range(5):
        t0 = time.perf_counter()
        for _ in range(200):
            L.kopek_fft_c2c_f64(pin, n, pout, npad)
        lat.append((time.perf_counter() - t0) / 200 * 1e6)
    t0 = time.perf_counter()
    for _ in range(200):
        fft.fft(x)
    py = (time.perf_counter() - t0) / 200 * 1e6
    print(f"fft() n_in={n:6d} -> {npad:6d}: C ABI {min(lat):7.1f} us (median {sorted(lat)[2]:7.1f}) | python mirror {py:7.1f} us")
