#!/usr/bin/env python3
"""Time the 1024-point kernels (KOPEK_K1W_VARIANT) on the headline shape and check them against K1 (variant -1).

Variants 0..13 are the three-pass kernel of fft_warp1024.cuh, which only tuning builds carry:
    python tools/tune.py build sweep=-DKOPEK_K1W_SWEEP
    KOPEK_TUNE_TAG=sweep python tools/sweep_k1w.py 1048576 -1,0,200        (200 = the production two-exchange kernel)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from kopek_b200 import StftPlan  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [-1] + list(range(14))
n = 1024
x = torch.randn(frames * n, device="cuda")
st = torch.cuda.current_stream().cuda_stream
ref = None
for window, output in ((1, 1), (0, 1)):
    for v in variants:
        os.environ["KOPEK_K1W_VARIANT"] = str(v)
        plan = StftPlan(n, n, window, output)
        out = torch.zeros((frames, plan.pitch), device="cuda", dtype=torch.complex64 if output == 0 else torch.float32)
        for _ in range(3):
            plan.exec_ptr(x.data_ptr(), x.numel(), out.data_ptr(), st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = int(os.environ.get("SWEEP_ITERS", "20"))
        e0.record()
        for _ in range(iters):
            plan.exec_ptr(x.data_ptr(), x.numel(), out.data_ptr(), st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        by = frames * (n * 4 + plan.bins * plan.elem_bytes)
        if v == variants[0]:
            ref = out.clone()
            err = 0.0
        else:
            a = torch.view_as_real(out) if output == 0 else out
            b = torch.view_as_real(ref) if output == 0 else ref
            if output == 3:
                a, b = 10 ** (a / 20), 10 ** (b / 20)
            err = ((a - b).abs().amax() / b.abs().amax()).item()
        print(f"win={window} out={output} variant={v:2d}: {ms:.4f} ms  {by / ms / 1e6:8.1f} GB/s  "
              f"{100 * by / ms / 1e6 / 6535.1:5.1f}% of 6535.1  maxdiff_vs_first={err:.2e}", flush=True)
        plan.close()
