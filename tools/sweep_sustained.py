#!/usr/bin/env python3
"""Burst AND sustained throughput of the batched kernels (hop = n, Hann |X|): the board's power cap pulls the SM clock
from 1.97 to ~1.6 GHz after ~50 ms of back-to-back launches, and a kernel that is HBM-bound at boost may be
issue-bound there.  Burst = 20 launches after 1.5 s of idle; sustained = 300 launches after 0.5 s of the same launch.

    python tools/sweep_sustained.py [sizes=1024,2048,4096] [bins=half|full] [hopdiv=1]
Run against tuning builds with tools/tune.py run tagA,tagB -- python tools/sweep_sustained.py ...
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from kopek_b200 import StftPlan  # noqa: E402

sizes = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1024, 2048, 4096]
full = len(sys.argv) > 2 and sys.argv[2] == "full"
hopdiv = int(sys.argv[3]) if len(sys.argv) > 3 else 1
PEAK = 6535.1
st = torch.cuda.current_stream().cuda_stream
try:
    import pynvml
    pynvml.nvmlInit()
    nv = pynvml.nvmlDeviceGetHandleByIndex(0)
    clock = lambda: pynvml.nvmlDeviceGetClockInfo(nv, pynvml.NVML_CLOCK_SM)
except Exception:
    clock = lambda: 0
for n in sizes:
    hop = n // hopdiv
    frames = (1 << 30) // n
    x = torch.randn(frames * n, device="cuda")
    fr = frames if hopdiv == 1 else min(frames, (frames * n - n) // hop + 1)
    ns = (fr - 1) * hop + n
    for window in (1, 0):
        plan = StftPlan(n, hop, window, 1, bins=1 if full else 0)
        out = torch.zeros((fr, plan.pitch), device="cuda")
        by = ns * 4 + fr * plan.bins * 4

        def run(k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(k):
                plan.exec_ptr(x.data_ptr(), ns, out.data_ptr(), st)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / k

        run(3)
        time.sleep(1.5)
        burst = run(20)
        t0 = time.time()
        while time.time() - t0 < 0.5:
            run(20)
        mhz = clock()
        sus = run(300)
        mhz2 = clock()
        print(f"n={n} hop={hop} win={window} {'full' if full else 'half'}: burst {burst:.4f} ms {100 * by / burst / 1e6 / PEAK:5.1f}%   "
              f"sustained {sus:.4f} ms {100 * by / sus / 1e6 / PEAK:5.1f}%  ({fr * n / sus / 1e6:6.1f} G samples/s, SM {mhz}/{mhz2} MHz)",
              flush=True)
        plan.close()
        time.sleep(1.0)
