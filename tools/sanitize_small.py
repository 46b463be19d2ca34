#!/usr/bin/env python3
"""Small end-to-end pass over every kernel for compute-sanitizer (one tool per gpurun call):
K1W (Hann mag, rect dB, bands), K1 (256/512/2048/4096, odd hop, full bins), K2 (small, chunked), K3 (2^16, 2^18), stream."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from kopek_b200 import (BINS_FULL, OUT_BANDS_LOG, OUT_BANDS_LOW, OUT_COMPLEX, OUT_DB, OUT_MAG, WINDOW_HANN,  # noqa: E402
                        WINDOW_RECT, LiveSpectrum, StftPlan, fft)

st = torch.cuda.current_stream().cuda_stream
rng = np.random.default_rng(0)


def run(n, hop, window, output, frames, bins=0):
    ns = (frames - 1) * hop + n
    x = torch.randn(ns, device="cuda")
    with StftPlan(n, hop, window, output, bins) as plan:
        out = torch.empty((frames, plan.pitch), device="cuda", dtype=torch.complex64 if output == 0 else torch.float32)
        assert plan.exec_ptr(x.data_ptr(), ns, out.data_ptr(), st) == frames
        torch.cuda.synchronize()
        assert bool(torch.isfinite(torch.view_as_real(out) if output == 0 else out).all()) or output == OUT_DB


for window, output in ((WINDOW_HANN, OUT_MAG), (WINDOW_RECT, OUT_DB), (WINDOW_HANN, OUT_COMPLEX),
                       (WINDOW_RECT, OUT_BANDS_LOG), (WINDOW_HANN, OUT_BANDS_LOW)):
    run(1024, 1024, window, output, 2000)          # K1W, more frames than resident warps
run(1024, 256, WINDOW_HANN, OUT_MAG, 700)          # K1W overlapped
run(1024, 333, WINDOW_HANN, OUT_MAG, 300)          # K1 scalar-load path
run(1024, 1024, WINDOW_RECT, OUT_MAG, 100, BINS_FULL)
for n in (256, 512, 2048, 4096):
    run(n, n, WINDOW_HANN, OUT_MAG, 600)
    run(n, n // 4, WINDOW_RECT, OUT_COMPLEX, 200)
for n_in in (1, 7, 1024, 5000, 20000):
    x = rng.standard_normal(n_in) + 1j * rng.standard_normal(n_in)
    fft.fft(x)
fft.fft_batched(rng.standard_normal((9, 300)) + 0j)
for lg in (16, 17, 18):
    a = torch.randn(1 << lg, dtype=torch.complex64, device="cuda")
    b = torch.empty_like(a)
    fft.fft_large_device(a.data_ptr(), b.data_ptr(), 1 << lg, 0, st)
    torch.cuda.synchronize()
live = LiveSpectrum(1024, WINDOW_HANN, OUT_MAG)
for c in (5, 100, 1024, 3000):
    live.push(rng.standard_normal(c).astype(np.float32))
live.close()
print("SANITIZE_RUN_OK")
