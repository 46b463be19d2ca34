#!/usr/bin/env python3
"""kopek_pcm16_scale_f32 (load_track_at_path, examples/graph/player.rs:81-102) on 2^28 values: 6 algorithmic bytes each."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from kopek_b200 import fft  # noqa: E402

n = 1 << 28
pcm = torch.randint(-32768, 32767, (n,), dtype=torch.int16, device="cuda")
out = torch.empty(n, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    fft.pcm16_scale_f32_device(pcm.data_ptr(), n, 0.00002, out.data_ptr(), stream=st)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    fft.pcm16_scale_f32_device(pcm.data_ptr(), n, 0.00002, out.data_ptr(), stream=st)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"pcm16_scale_f32 2^28 values: {ms:.4f} ms = {n * 6 / ms / 1e6:.0f} GB/s = {100 * n * 6 / ms / 1e6 / 6535.1:.1f} % of the roofline")
a = torch.empty(n * 6 // 8, dtype=torch.float32, device="cuda")        # a copy of the same total bytes (1:1 read:write)
b = torch.empty_like(a)
for _ in range(3):
    b.copy_(a)
e0.record()
for _ in range(10):
    b.copy_(a)
e1.record()
torch.cuda.synchronize()
print(f"torch copy of the same bytes: {e0.elapsed_time(e1) / 10:.4f} ms")
