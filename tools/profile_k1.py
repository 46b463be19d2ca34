#!/usr/bin/env python3
"""Minimal K1 driver for ncu: a few launches of one plan on the headline shape, nothing else.

    python tools/profile_k1.py [--n 1024] [--hop 1024] [--frames 1048576] [--window 1] [--output 1] [--iters 5]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from kopek_b200 import StftPlan  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1024)
ap.add_argument("--hop", type=int, default=0)
ap.add_argument("--frames", type=int, default=1 << 20)
ap.add_argument("--window", type=int, default=1)
ap.add_argument("--output", type=int, default=1)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--pitch", type=int, default=0)
ap.add_argument("--bulk", action="store_true")
ap.add_argument("--full", action="store_true", help="all n bins per row")
a = ap.parse_args()
hop = a.hop or a.n
ns = (a.frames - 1) * hop + a.n
x = torch.randn(ns, device="cuda")
plan = StftPlan(a.n, hop, a.window, a.output, bins=1 if a.full else 0, out_pitch=a.pitch, bulk_store=a.bulk)
out = torch.empty((a.frames, plan.pitch), device="cuda", dtype=torch.complex64 if a.output == 0 else torch.float32)
st = torch.cuda.current_stream().cuda_stream
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(a.iters):
    if i == a.iters - 1:
        ev0.record()
    plan.exec_ptr(x.data_ptr(), ns, out.data_ptr(), st)
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1)
by = a.frames * (a.n * 4 + plan.bins * plan.elem_bytes)
print(f"n={a.n} hop={hop} frames={a.frames} last launch {ms:.4f} ms  {by / ms / 1e6:.1f} GB/s algorithmic")
