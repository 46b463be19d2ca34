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

# the K2 kernel alone (device-resident buffers, back-to-back launches on one stream: launch rate or kernel time)
import torch  # noqa: E402  (device memory only)

for n in (256, 1024, 2048, 4096):
    d_in = torch.randn(2 * n, device="cuda", dtype=torch.float64)
    d_out = torch.empty(2 * n, device="cuda", dtype=torch.float64)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(10):
        L.kopek_fft_c2c_f64_device(d_in.data_ptr(), n, 1, d_out.data_ptr(), 0, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(500):
        L.kopek_fft_c2c_f64_device(d_in.data_ptr(), n, 1, d_out.data_ptr(), 0, st)
    e1.record()
    torch.cuda.synchronize()
    print(f"K2 kernel, device-resident n={n:5d}: {e0.elapsed_time(e1) / 500 * 1e3:6.2f} us per back-to-back launch")

# the live window (kopek_stream_push): one UI frame = push the new samples, get the spectrum of the latest n
from kopek_b200 import OUT_MAG, LiveSpectrum  # noqa: E402

for n, count in ((1024, 1024), (1024, 480), (4096, 512)):
    live = LiveSpectrum(n, 1, OUT_MAG)
    new = rng.standard_normal(count).astype(np.float32)
    for _ in range(20):
        live.push(new)
    lat = []
    for _ in range(5):
        t0 = time.perf_counter()
        for _ in range(200):
            live.push(new)
        lat.append((time.perf_counter() - t0) / 200 * 1e6)
    live.close()
    print(f"stream push n={n:5d} count={count:5d} (Hann |X|): {min(lat):7.1f} us (median {sorted(lat)[2]:7.1f})")
