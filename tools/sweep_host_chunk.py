"""e2e leg of bench.py (kopek_stft_exec_host, pinned host buffers) against the chunk size of the H2D/kernel/D2H
pipeline.  KOPEK_HOST_CHUNK_MB is read by the library on every call."""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, ".")
from kopek_b200 import OUT_MAG, WINDOW_HANN, StftPlan, _lib  # noqa: E402

N, BINS = 1024, 513
lib = _lib.load()
for frames in (1 << 18, 1 << 16):
    in_bytes, out_bytes = frames * N * 4, frames * BINS * 4
    h_in, h_out = C.c_void_p(), C.c_void_p()
    _lib.check(lib.kopek_host_alloc(C.byref(h_in), in_bytes))
    _lib.check(lib.kopek_host_alloc(C.byref(h_out), out_bytes))
    x = torch.randn(frames * N)
    C.memmove(h_in.value, x.data_ptr(), in_bytes)
    for mb in (32, 16, 8, 4, 2, 1):
        os.environ["KOPEK_HOST_CHUNK_MB"] = str(mb)
        plan = StftPlan(N, N, WINDOW_HANN, OUT_MAG, device=0)
        plan.exec_host_ptr(h_in.value, frames * N, h_out.value)
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                plan.exec_host_ptr(h_in.value, frames * N, h_out.value)
            best = min(best, (time.perf_counter() - t0) / 5)
        print(f"frames={frames:7d} chunk={mb:3d} MiB: {best * 1e3:7.3f} ms  {frames * N / best / 1e9:6.2f} G samples/s  "
              f"H2D {in_bytes / best / 1e9:5.1f} GB/s  launches {plan.last_launches}")
        plan.close()
    lib.kopek_host_free(h_in)
    lib.kopek_host_free(h_out)
