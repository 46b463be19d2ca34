#!/usr/bin/env python3
"""Minimal K3 driver for ncu: a few 2^24-point four-step transforms."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from kopek_b200 import fft  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << lg
a = torch.randn(n, dtype=torch.complex64, device="cuda")
b = torch.empty_like(a)
st = torch.cuda.current_stream().cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(4):
    if i == 3:
        e0.record()
    fft.fft_large_device(a.data_ptr(), b.data_ptr(), n, 0, st)
e1.record()
torch.cuda.synchronize()
print(f"2^{lg}: last transform {e0.elapsed_time(e1):.4f} ms")
