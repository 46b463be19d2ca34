#!/usr/bin/env python3
"""Time the four-step transform at 2^16 .. 2^24 (20 launches after 3 warm-ups) and check it against torch.fft."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from kopek_b200 import fft  # noqa: E402

st = torch.cuda.current_stream().cuda_stream
for lg in [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else (16, 18, 20, 22, 24):
    n = 1 << lg
    a = torch.randn(n, dtype=torch.complex64, device="cuda")
    b = torch.empty_like(a)
    for _ in range(3):
        fft.fft_large_device(a.data_ptr(), b.data_ptr(), n, 0, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fft.fft_large_device(a.data_ptr(), b.data_ptr(), n, 0, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    ref = torch.fft.fft(a.to(torch.complex128))
    err = ((b.to(torch.complex128) - ref).abs().max() / ref.abs().max()).item()
    print(f"2^{lg}: {ms:.4f} ms  {2 * n * 8 / ms / 1e6:7.1f} GB/s algorithmic  {100 * 2 * n * 8 / ms / 1e6 / 6535.1:5.1f}%  "
          f"max rel err vs f64 {err:.2e}", flush=True)
