#!/usr/bin/env python3
"""Three calls of K3 and of cuFFT C2C per size, nothing else: the program to put under
`ncu --metrics gpu__time_duration.sum` when the question is which PASS of a size is the slow one.

    python tools/k3_probe.py 20,21
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from kopek_b200 import fft  # noqa: E402

st = torch.cuda.current_stream().cuda_stream
for lg in [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "20,21").split(",")]:
    n = 1 << lg
    a = torch.randn(n, dtype=torch.complex64, device="cuda")
    b = torch.empty_like(a)
    for _ in range(3):
        fft.fft_large_device(a.data_ptr(), b.data_ptr(), n, 0, st)
    torch.cuda.synchronize()
    for _ in range(3):
        c = torch.fft.fft(a)
    torch.cuda.synchronize()
