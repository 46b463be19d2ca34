#!/usr/bin/env python3
"""K3 (four-step) timing and error against an f64 transform, 2^16 .. 2^24 points, with cuFFT C2C f32 beside it.

    python tools/bench_k3.py [log2 sizes = 16,18,20,22,23,24]
KOPEK_K3_KEEP_MB (read by the library) selects how much of the inter-pass workspace is asked to stay in L2.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from kopek_b200 import fft  # noqa: E402

sizes = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [16, 18, 20, 22, 23, 24]
st = torch.cuda.current_stream().cuda_stream
PEAK = 6535.1


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def cold_timed(fn, iters=10):
    """Each call timed on its own after a 512 MB write has pushed the previous call's data out of L2 (the eager and the
    graph figures re-run on the same buffers: for n <= 2^22 input and output are then L2-resident)."""
    flush = torch.empty(128 << 20, dtype=torch.float32, device="cuda")
    tot = 0.0
    for _ in range(iters):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def graph_timed(fn, calls=20, reps=5):
    """The same calls captured ONCE into a CUDA graph and replayed: device time per call with no host work between
    launches (the eager figure includes whatever the host needs per call when that exceeds the kernels' time)."""
    try:
        side = torch.cuda.Stream()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(calls):
                fn(torch.cuda.current_stream().cuda_stream)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (reps * calls)
    except Exception as ex:                                   # capture not possible: say so, keep the eager figures
        torch.cuda.synchronize()
        print(f"   (graph capture failed: {type(ex).__name__}: {str(ex)[:120]})")
        return float("nan")


print(f"KOPEK_K3_KEEP_MB={os.environ.get('KOPEK_K3_KEEP_MB', '(default)')} lib={os.environ.get('KOPEK_TUNE_TAG', 'production')}")
for lg in sizes:
    n = 1 << lg
    a = torch.randn(n, dtype=torch.complex64, device="cuda")
    a[12345 % n] += 100.0
    b = torch.empty_like(a)
    ms = timed(lambda: fft.fft_large_device(a.data_ptr(), b.data_ptr(), n, 0, st))
    lib = timed(lambda: torch.fft.fft(a))
    want = torch.fft.fft(a.to(torch.complex128))
    err = ((b.to(torch.complex128) - want).abs().max() / want.abs().max()).item()
    gbs = 2 * n * 8 / ms / 1e6
    c_ms = cold_timed(lambda: fft.fft_large_device(a.data_ptr(), b.data_ptr(), n, 0, st))
    c_lib = cold_timed(lambda: torch.fft.fft(a))
    g_ms = graph_timed(lambda s_: fft.fft_large_device(a.data_ptr(), b.data_ptr(), n, 0, s_))
    g_lib = graph_timed(lambda s_: torch.fft.fft(a))
    print(f"2^{lg}: {ms:.4f} ms  {gbs:7.1f} GB/s algorithmic {100 * gbs / PEAK:5.1f}%   cuFFT {lib:.4f} ms ({lib / ms:.2f}x)   "
          f"max rel err vs f64 {err:.2e}   | in a CUDA graph: {g_ms:.4f} ms, cuFFT {g_lib:.4f} ms ({g_lib / g_ms:.2f}x) | cold L2, single call: {c_ms:.4f} ms, cuFFT {c_lib:.4f} ms ({c_lib / c_ms:.2f}x)", flush=True)
