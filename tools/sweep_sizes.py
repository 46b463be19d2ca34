#!/usr/bin/env python3
"""Time the batched kernel of every frame size (hop = n, Hann |X|) against the generic K1 (KOPEK_K1W_VARIANT=-1)
and report the difference between the two results.

    python tools/sweep_sizes.py [sizes=256,512,1024,2048,4096] [variants=-1,0]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from kopek_b200 import StftPlan  # noqa: E402

sizes = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [256, 512, 1024, 2048, 4096]
variants = sys.argv[2].split(",") if len(sys.argv) > 2 else ["-1", ""]
st = torch.cuda.current_stream().cuda_stream
for n in sizes:
    frames = (1 << 30) // n
    x = torch.randn(frames * n, device="cuda")
    for window, output, hop in ((1, 1, n), (0, 1, n), (1, 3, n // 4)):
        fr = frames if hop == n else (frames * n - n) // hop + 1
        if hop != n:
            fr = min(fr, frames)      # keep the output the same size
        ns = (fr - 1) * hop + n
        ref = None
        for v in variants:
            if v == "":
                os.environ.pop("KOPEK_K1W_VARIANT", None)
            else:
                os.environ["KOPEK_K1W_VARIANT"] = v
            plan = StftPlan(n, hop, window, output)
            out = torch.zeros((fr, plan.pitch), device="cuda")
            for _ in range(3):
                plan.exec_ptr(x.data_ptr(), ns, out.data_ptr(), st)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                plan.exec_ptr(x.data_ptr(), ns, out.data_ptr(), st)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            by = ns * 4 + fr * plan.bins * 4
            if ref is None:
                ref, err = out.clone(), 0.0
            else:
                a, b = (10 ** (out / 20), 10 ** (ref / 20)) if output == 3 else (out, ref)
                err = ((a - b).abs().amax() / b.abs().amax()).item()
            print(f"n={n} hop={hop} win={window} out={output} variant={v or 'default':>7}: {ms:.4f} ms {by / ms / 1e6:8.1f} GB/s "
                  f"{100 * by / ms / 1e6 / 6535.1:5.1f}%  {fr * n / ms / 1e6:7.1f} G samples/s  diff={err:.2e}", flush=True)
            plan.close()
