#!/usr/bin/env python3
"""Burst vs sustained HBM copy bandwidth with the method of MEASURED_PEAKS.json (torch b.copy_(a), 1 Gi bf16, read+write bytes)."""
import torch
a = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda").normal_()
b = torch.empty_like(a)
by = 2 * a.numel() * 2
def run(iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        b.copy_(a)
    e1.record(); torch.cuda.synchronize()
    return by * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9
for _ in range(3): b.copy_(a)
torch.cuda.synchronize()
best = max(run(1) for _ in range(10))
print(f"copy best-of-10 single: {best:.1f} GB/s")
print(f"copy 20 back to back:   {run(20):.1f} GB/s")
print(f"copy 400 back to back:  {run(400):.1f} GB/s")
print(f"copy 400 again:         {run(400):.1f} GB/s")
