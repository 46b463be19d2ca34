#!/usr/bin/env python3
"""bench.py -- headline benchmark of the kopek FFT+magnitude hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json metric "complex samples/sec batched 1024-pt FFT+mag", SURVEY.md 8d
config M): per GPU 2^20 non-overlapping frames of 1024 real f32 samples (sine of
20*(1 + f mod 1000) Hz at 48 kHz + 0.1*N(0,1)), Hann window, 513 f32 |X| per frame.
A "step" is one pass of the fused K1 kernel over that batch.  N > 1 runs one
process per GPU (torchrun), each rank transforms its own 2^20 frames (weak
scaling, no data-path collective).  torch is used for device memory, streams,
events and torch.distributed only; the transform is libkopek_cuda.so.

One JSON line on stdout (rank 0):
  value         whole-job samples/s with inputs resident in HBM (K timed steps, CUDA events, max over ranks)
  e2e           the same metric through the host-buffer C-ABI call (pinned host buffers, H2D + kernel + D2H inside
                the timed region), with the box's plain-copy ceiling measured beside it
  roofline      the K1 kernel against MEASURED_PEAKS.json (burst over the K steps, sustained over 300 more)
  parity        the headline output against the f64 and the f32 oracle on a slice of the very buffers that were timed
  configs       every BASELINE.json config (cfg0..cfg4): ms, fraction of the roofline, parity on an oracle-checked
                slice; cfg2 strong-scaled over the ranks, cfg4 one full-size channel per GPU
  gather        (N > 1) ONE spectrogram on rank 0: NCCL gather and the fused peer-direct bulk-store form, both checked
                bit-identical to the 1-GPU result
  dist_fft      (N in 2, 4, 8) one transform of N * 2^22 points over all GPUs against an f64 transform
  mgpu          (N > 1) the C ABI's single-process multi-GPU entry points driven by rank 0 alone
  cpu_baseline  the oracle port of src/fft.rs on all host cores; cpu_baseline_1core the same on one (rank 0, N = 1)
`--impl reference` times that CPU port alone with all host threads.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import math
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FFT = 1024
BINS = N_FFT // 2 + 1
FRAMES_PER_GPU = 1 << 20
E2E_FRAMES = 1 << 20
ALGO_BYTES_PER_FRAME = N_FFT * 4 + BINS * 4          # 6148: samples read once + magnitudes written once
METRIC = "complex samples/sec batched 1024-pt FFT+mag"
UNIT = "samples/s"
WORKLOAD = "M: 2^20 frames x 1024-pt real f32 per GPU, Hann window, |X| (513 f32/frame), non-overlapping"
HEADLINE_KERNEL_SOURCES = ("fft_warp1024b.cuh", "lane_common.cuh", "fft_batched.cuh", "k1w_inst.cu")
L2_BYTES = 126 << 20
# Static instruction mix of the frame loop of the kernels the configs run on, per LANE and frame (tools/loop_mix.py on the
# shipped build; tests/test_kernel_models.py::test_bench_static_mix_matches_the_library keeps them in step): the
# overlapped STFT configs transform every sample N/hop times, so their roofs are FP32 throughput and instruction issue,
# not HBM.  flops: FFMA and the packed FADD2 count 2, FADD / FMUL 1.
KERNEL_MIX = {
    "k1w_1024b<MAG,WIN>": {"symbol": "k1w_1024b_kernelILi1ELb1ELi4ELi3ELb0ELb0ELb0ELb1E", "lanes": 32, "loop": 789,
                           "FFMA": 207, "FMUL": 112, "FADD": 273, "FADD2": 0},
    "k1m<4,DB,cos-sum>": {"symbol": "k1m_kernelILi4ELi3ELi2ELi6ELi0ELb0ELb1E", "lanes": 64, "loop": 828,
                          "FFMA": 242, "FMUL": 144, "FADD": 141, "FADD2": 108},
    "k1m<8,MAG,cos-sum>": {"symbol": "k1m_kernelILi8ELi1ELi2ELi3ELi0ELb0ELb1E", "lanes": 128, "loop": 897,
                           "FFMA": 249, "FMUL": 180, "FADD": 135, "FADD2": 124},
}
SM_COUNT, FP32_LANES_PER_SM, SCHEDULERS_PER_SM = 148, 128, 4


def compute_roofs(mix_key: str, frames: int, ms: float, sm_mhz: float):
    """FP32 and instruction-issue roofs of one launch at the SM clock `sm_mhz` (the board's maximum: an upper bound)."""
    m = KERNEL_MIX[mix_key]
    flops = m["lanes"] * (2 * m["FFMA"] + m["FMUL"] + m["FADD"] + 2 * m["FADD2"])
    warp_instr = m["lanes"] // 32 * m["loop"]
    fp32_peak = SM_COUNT * FP32_LANES_PER_SM * 2 * sm_mhz * 1e6
    issue_peak = SM_COUNT * SCHEDULERS_PER_SM * sm_mhz * 1e6
    t = ms * 1e-3
    return {"kernel_mix": mix_key, "flops_per_frame": flops, "warp_instructions_per_frame": warp_instr,
            "fp32": {"achieved_tflops": frames * flops / t / 1e12, "peak_tflops": fp32_peak / 1e12,
                     "frac": frames * flops / t / fp32_peak},
            "issue": {"achieved_ginstr_s": frames * warp_instr / t / 1e9, "peak_ginstr_s": issue_peak / 1e9,
                      "frac": frames * warp_instr / t / issue_peak},
            "clock_mhz": sm_mhz, "note": "static SASS mix of the frame loop (profiles/r2_sass_summary.md); peaks at the "
                                         "board's maximum SM clock"}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def headline_kernel_fingerprint() -> str:
    """Hash of the sources the headline kernel is compiled from: an ncu capture is only quoted for the build it saw."""
    h = hashlib.sha256()
    for name in HEADLINE_KERNEL_SOURCES:
        with open(os.path.join(ROOT, "kopek_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def _ncu_traffic():
    """(dram bytes per K1 launch, note) from the committed ncu capture -- refused when it was taken from another build."""
    try:
        with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as f:
            rec = json.load(f)
    except Exception:
        return None, "no committed ncu capture"
    if rec.get("kernel_src_sha") != headline_kernel_fingerprint():
        return None, (f"profiles/k1_traffic.json was captured from another build of the kernel "
                      f"({rec.get('kernel', '?')}, sources {rec.get('kernel_src_sha', 'unrecorded')}): not quoted")
    return rec.get("dram_bytes_per_launch"), f"ncu --set full, {rec.get('kernel')}, {rec.get('source')}"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def _physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def bind_near_gpu(index: int):
    """Pin this rank's threads to the CPUs NVML reports as local to its GPU, BEFORE the pinned host buffers of the
    e2e leg are allocated, so that they come from the GPU's own NUMA node.  Returns the number of CPUs, None if unknown
    (a VM that exposes one NUMA node reports all of its CPUs for every GPU: the call is then a no-op)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def claim_stdout():
    """stdout carries ONE JSON line.  Libraries print there too (NCCL writes its version banner to stdout when
    NCCL_DEBUG is set), so file descriptor 1 is pointed at stderr for the rest of the process and the line is written
    to the descriptor stdout had when the process started."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


# ------------------------------------------------------------------------------------------------ CPU legs ----
def cpu_port_rate(frames: int, threads: int):
    """Time the oracle port (restatement of src/fft.rs) on `frames` frames of the workload."""
    import oracle  # CPU baseline leg: the one place bench.py executes oracle/
    from kopek_b200 import signals
    x = signals.sine_frames(frames, N_FFT)
    t0 = time.perf_counter()
    oracle.stft_mag(x, N_FFT, N_FFT, oracle.WIN_HANN, output=oracle.OUT_MAG, threads=threads)
    dt = time.perf_counter() - t0
    return frames * N_FFT / dt, dt


def cpu_baseline(target_s: float = 12.0, cores: int | None = None):
    import oracle
    oracle.build()
    cores = oracle.max_threads() if cores is None else cores
    probe = 256 * cores
    rate, _ = cpu_port_rate(probe, cores)
    frames = int(max(probe, min(rate * target_s / N_FFT, 4_000_000)))
    rate, dt = cpu_port_rate(frames, cores)
    return {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{frames} frames of the workload ({dt:.1f} s): oracle/ref_fft.c, f64 recursion of "
                      f"src/fft.rs + Hann + norm(), {cores} pthread{'s' if cores > 1 else ''} over independent frames"}


def run_reference(args):
    """--impl reference: the reference algorithm's CPU port, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    out = claim_stdout()
    import oracle
    oracle.build()
    cores = oracle.max_threads()
    frames = 4096 * cores       # ~0.3 s per step on 16 cores: long enough that thread start-up does not understate the port
    for _ in range(args.warmup):
        cpu_port_rate(min(frames, 64 * cores), cores)
    from kopek_b200 import signals
    x = signals.sine_frames(frames, N_FFT)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.stft_mag(x, N_FFT, N_FFT, oracle.WIN_HANN, output=oracle.OUT_MAG, threads=cores)
    dt = time.perf_counter() - t0
    value = frames * N_FFT * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_frames_per_step": frames,
                   "note": "CPU port of src/fft.rs (Rust reference cannot be built in this image); "
                           "window applied on the host, norm() magnitudes"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{frames} frames per step x {args.steps} steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=out, flush=True)
    return 0


# ------------------------------------------------------------------------------------------- parity helpers ----
def parity_against_oracle(got, samples, n, hop, window, output, full=False):
    """got: [frames, n/2+1] of the GPU path (|X|, |X|^2 or dB) for the frames of `samples`; compared in the linear
    domain with the f64 and the f32 instantiation of the restated recursion (north_star: 1e-5 * log2 N per bin relative
    to the frame's largest bin; peak bins equal, a near-tie closer than the tolerance is not a mismatch)."""
    import numpy as np

    import oracle
    win = {0: oracle.WIN_RECT, 1: oracle.WIN_HANN, 2: oracle.WIN_HAMMING}[window]
    got = np.asarray(got, dtype=np.float64)
    if output == 3:
        got = 10.0 ** (got / 20.0)
    elif output == 2:
        got = np.sqrt(got)
    tol = 1e-5 * math.log2(n)
    res = {"frames_checked": int(got.shape[0]), "tolerance": tol}
    ref64 = None
    for prec in (64, 32):
        ref = np.abs(oracle.stft_complex(samples, n, hop, win, bins=n if full else None, precision=prec,
                                         threads=oracle.max_threads()))
        den = ref.max(axis=1, keepdims=True)
        den[den == 0] = 1.0
        res[f"max_rel_err_vs_f{prec}_oracle"] = float((np.abs(got - ref) / den).max())
        if prec == 64:
            ref64 = ref
    half = n // 2 + 1                                         # peaks over bins 0..N/2 (the mirror repeats them)
    pg, pr = got[:, :half].argmax(axis=1), ref64[:, :half].argmax(axis=1)
    bad = 0
    for f in np.nonzero(pg != pr)[0]:
        top = ref64[f, pr[f]]
        if abs(ref64[f, pg[f]] - top) > 2 * tol * top:
            bad += 1
    res["peak_bin_mismatches"] = bad
    res["ok"] = bool(res["max_rel_err_vs_f64_oracle"] <= tol and res["max_rel_err_vs_f32_oracle"] <= tol and bad == 0)
    return res


def block_checksums(torch, t, piece: int = 1 << 26):
    """Two 64-bit checksums of a tensor's bits (the second order-sensitive), accumulated piece by piece so that an
    11 GB block costs 1.5 GB of temporaries: equal on both sides of a transfer <=> the same bits arrived."""
    v = t.contiguous().view(torch.int32).reshape(-1)
    a = b = 0
    for off in range(0, v.numel(), piece):
        w = v[off:off + piece].to(torch.int64)
        idx = torch.arange(off, off + w.numel(), device=w.device, dtype=torch.int64)
        a = (a + int(w.sum().item())) & 0xFFFFFFFFFFFFFFFF
        b = (b + int((w * ((idx % 65521) + 1)).sum().item())) & 0xFFFFFFFFFFFFFFFF
        del w, idx
    return a, b


# ----------------------------------------------------------------------------------------------- the bench ----
class Ctx:
    pass


def time_launches(torch, fn, reps, warm=3):
    """Mean device time of `fn(i)` over `reps` back-to-back calls after `warm` untimed ones (CUDA events on the
    current stream)."""
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(warm + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def time_graph_replays(torch, fn, calls=20, reps=5):
    """`fn(stream)` captured `calls` times into ONE CUDA graph and replayed `reps` times: device time per call with no
    host work between the launches.  None when the capture is not possible (never fatal to the bench line)."""
    try:
        side = torch.cuda.Stream()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
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
    except Exception:
        torch.cuda.synchronize()
        return None


def max_over_ranks(c, value: float) -> float:
    if c.world == 1:
        return value
    t = c.torch.tensor([value], device=c.device, dtype=c.torch.float64)
    c.dist.all_reduce(t, op=c.dist.ReduceOp.MAX)
    return float(t.item())


def copy_ceiling(c, in_bytes: int, out_bytes: int, reps: int = 3):
    """What the box's PCIe/host-memory path delivers with NO kernel: every rank, at the same time, copies `in_bytes`
    host->device and `out_bytes` device->host from pinned buffers on two streams (one cudaMemcpyAsync each), like the
    e2e leg's traffic.  Returns whole-job samples/s equivalent (4 bytes in per sample) -- the ceiling of e2e."""
    torch = c.torch
    h_in = torch.empty(in_bytes, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(out_bytes, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(in_bytes, dtype=torch.uint8, device=c.device)
    d_out = torch.empty(out_bytes, dtype=torch.uint8, device=c.device)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    chunk = 16 << 20

    def whole():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    def chunked():                         # the shape of the e2e leg's traffic: 16 MiB pieces in both directions
        for off in range(0, in_bytes, chunk):
            with torch.cuda.stream(s1):
                d_in[off:off + chunk].copy_(h_in[off:off + chunk], non_blocking=True)
            o0, o1 = off * out_bytes // in_bytes, min(out_bytes, (off + chunk) * out_bytes // in_bytes)
            with torch.cuda.stream(s2):
                h_out[o0:o1].copy_(d_out[o0:o1], non_blocking=True)

    best = 0.0
    for once in (whole, chunked):
        once()
        s1.synchronize()
        s2.synchronize()
        c.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            once()
            s1.synchronize()
            s2.synchronize()
        c.barrier()
        dt = max_over_ranks(c, time.perf_counter() - t0)
        best = max(best, c.world * (in_bytes / 4) * reps / dt)
    return best


def bench_config_cfg0(c):
    """configs[0]: one 1024-point fft() drop-in call on the 440 Hz / 44.1 kHz oscillator sine (latency, bit-exact)."""
    import numpy as np

    import oracle
    from kopek_b200 import fft
    s = oracle.osc_sine(44100.0, 440.0, 1024)
    z = oracle.marshal(s)
    for _ in range(20):
        X = fft.fft(z)
    t0 = time.perf_counter()
    reps = 200
    for _ in range(reps):
        X = fft.fft(z)
    us = (time.perf_counter() - t0) / reps * 1e6
    ref = oracle.fft(z)
    t0 = time.perf_counter()
    for _ in range(20):
        oracle.fft(z)
    cpu_us = (time.perf_counter() - t0) / 20 * 1e6
    return {"workload": "cfg0: single 1024-pt fft() of Oscillator::sine 440 Hz @ 44.1 kHz, host buffers in and out",
            "kernel": "k2_c2c_f64_chunk (f64, bit-exact radix-2 DIT)", "us_per_call": us, "cpu_port_us_per_call": cpu_us,
            "bit_exact_vs_oracle": bool(np.array_equal(X.view(np.uint64), ref.view(np.uint64))),
            "peak_bin": int(np.abs(X[:513]).argmax()), "peak_mag": float(np.abs(X[:513]).max())}


def bench_stft_config(c, name, n, hop, window, output, samples_dev, frames_expected, kernel, workload, check_frames,
                      reps=20, copies=1, full=False):
    """Times one plan over device-resident samples (`copies` > 1 rotates through as many input/output buffer pairs so
    that a working set smaller than L2 is not re-read from it) and checks a slice against the oracle."""
    torch = c.torch
    from kopek_b200 import StftPlan
    bins = n if full else n // 2 + 1
    stream = torch.cuda.current_stream().cuda_stream
    with StftPlan(n, hop, window, output, bins=1 if full else 0, device=c.local) as plan:
        ins = [samples_dev] + [samples_dev.clone() for _ in range(copies - 1)]
        outs = [torch.empty((frames_expected, bins), device=c.device, dtype=torch.float32) for _ in range(copies)]

        def fn(i):
            plan.exec_ptr(ins[i % copies].data_ptr(), ins[i % copies].numel(), outs[i % copies].data_ptr(), stream)

        ms = time_launches(torch, fn, reps)
        assert plan.frames(samples_dev.numel()) == frames_expected
        ms = max_over_ranks(c, ms)
        res = {"workload": workload, "kernel": kernel, "ms": ms, "frames": frames_expected, "n": n, "hop": hop,
               "launches_timed": reps}
        in_bytes = samples_dev.numel() * 4
        out_bytes = frames_expected * bins * 4
        res["algorithmic_bytes"] = in_bytes + out_bytes
        res["achieved_gbs"] = (in_bytes + out_bytes) / (ms * 1e-3) / 1e9
        res["frac"] = res["achieved_gbs"] / c.peak
        res["transformed_samples_per_s"] = frames_expected * n / (ms * 1e-3)
        ws = copies * (in_bytes + out_bytes)
        res["l2"] = (f"working set {ws / 1e6:.0f} MB over {copies} rotating buffer pairs > 126 MB L2" if copies > 1
                     else f"working set {ws / 1e6:.0f} MB > 126 MB L2")
        if c.rank == 0 and check_frames:
            k = min(check_frames, frames_expected)
            ns = (k - 1) * hop + n
            res["parity"] = parity_against_oracle(outs[0][:k].cpu().numpy(), samples_dev[:ns].cpu().numpy(), n, hop,
                                                  window, output, full=full)
        del ins, outs
    return res


def run_ours(args):
    import ctypes as C

    import numpy as np
    import torch
    import torch.distributed as dist

    from kopek_b200 import OUT_DB, OUT_MAG, WINDOW_HANN, WINDOW_RECT, StftPlan, _lib, fft, signals

    c = Ctx()
    c.torch, c.dist = torch, dist
    c.rank = rank = int(os.environ.get("RANK", "0"))
    c.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    c.local = local = int(os.environ.get("LOCAL_RANK", "0"))
    out_stream = claim_stdout()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: kopek_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    c.device = device = torch.device("cuda", local)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        cpu_group = dist.new_group(backend="gloo")       # host-side waits that keep the GPUs idle (mgpu leg)
    c.peak, peak_src = _peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    c.barrier = barrier

    all_cpus = os.sched_getaffinity(0)
    near = bind_near_gpu(_physical_gpu_index(local))
    frames = args.frames
    plan = StftPlan(N_FFT, N_FFT, WINDOW_HANN, OUT_MAG, device=local)
    stream = torch.cuda.current_stream().cuda_stream

    def synth(count_frames, n, first_frame, **kw):
        x = torch.empty(count_frames * n, device=device, dtype=torch.float32)
        fft.synth_frames_device(x.data_ptr(), count_frames, n, 48000.0, first_frame=first_frame, device=local,
                                stream=stream, **kw)
        torch.cuda.synchronize()
        return x

    # Workload M generated on the device by the library's own oscillator/noise sources (kopek_synth_frames):
    # frame f = Oscillator::sine at 20*(1 + f mod 1000) Hz / 48 kHz from phase 0, + 0.1 * N(0,1), seed 1234
    m_kw = dict(wave=fft.WAVE_SINE, freq0=20.0, freq_step=20.0, freq_mod=1000, noise=fft.NOISE_WHITE, noise_gain=0.1,
                seed=1234)
    x = synth(frames, N_FFT, rank * frames, **m_kw)
    out = torch.empty((frames, BINS), device=device, dtype=torch.float32)

    def step():
        plan.exec_ptr(x.data_ptr(), x.numel(), out.data_ptr(), stream)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(_physical_gpu_index(local))
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    launches = 0
    ev[0].record()
    for i in range(args.steps):
        step()
        launches += plan.last_launches
        ev[i + 1].record()
    barrier()
    clocks = sampler.stop()
    total_ms = ev[0].elapsed_time(ev[args.steps])
    # sustained regime: the same launch back to back for ~0.4 s (the board reaches its power cap and the SM
    # clock settles below boost); reported beside the K-step number, never instead of it
    sustained = None
    if args.sustained_steps > 0:
        s2 = ClockSampler(_physical_gpu_index(local))
        s2.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.sustained_steps):
            step()
        e1.record()
        barrier()
        c2 = s2.stop()
        sustained = {"steps": args.sustained_steps, "ms_per_step": e0.elapsed_time(e1) / args.sustained_steps,
                     "clocks": c2}
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    max_ms = max_over_ranks(c, total_ms)
    value = world * frames * N_FFT * args.steps / (max_ms * 1e-3)

    # ---- parity of the very buffers that were timed (first frames of this rank's batch), against both oracles
    parity = None
    if rank == 0 and not args.no_parity:
        import oracle
        oracle.build()
        k = 256
        parity = parity_against_oracle(out[:k].cpu().numpy(), x[:k * N_FFT].cpu().numpy(), N_FFT, N_FFT, 1, 1)

    # ---- end to end: pinned host buffers -> kopek_stft_exec_host -> pinned host spectra, full headline batch
    lib = _lib.load()
    ef = min(args.e2e_frames, frames)
    h_in, h_out = C.c_void_p(), C.c_void_p()
    while True:
        in_bytes, out_bytes = ef * N_FFT * 4, ef * BINS * 4
        ok = lib.kopek_host_alloc(C.byref(h_in), in_bytes) == 0
        ok = ok and lib.kopek_host_alloc(C.byref(h_out), out_bytes) == 0
        if ok or ef <= (1 << 16):
            break
        lib.kopek_host_free(h_in)       # a small VM may not pin 6.4 GB: halve and say so (frames_per_step reports it)
        h_in = C.c_void_p()
        ef //= 2
    if not ok:
        raise SystemExit("cannot allocate pinned host buffers for the e2e leg: " + lib.kopek_last_error().decode())
    step_b = 1 << 16
    for f0 in range(0, ef, step_b):              # device -> pinned host in slices (no second 4 GiB pageable copy)
        nfr = min(step_b, ef - f0)
        piece = x[f0 * N_FFT:(f0 + nfr) * N_FFT].cpu()          # keep the host copy alive across the memmove
        C.memmove(h_in.value + f0 * N_FFT * 4, piece.data_ptr(), nfr * N_FFT * 4)
        del piece
    e2e_steps = max(3, min(args.steps, 5))
    plan.exec_host_ptr(h_in.value, ef * N_FFT, h_out.value)          # warm-up (allocates the chunk slots)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        plan.exec_host_ptr(h_in.value, ef * N_FFT, h_out.value)
    e2e_launches = plan.last_launches
    barrier()
    e2e_s = max_over_ranks(c, time.perf_counter() - t0)
    e2e_value = world * ef * N_FFT * e2e_steps / e2e_s
    # the spectra that came back over PCIe equal the device-resident ones
    same = True
    for f0 in range(0, ef, step_b):
        nfr = min(step_b, ef - f0)
        chk = torch.empty(nfr * BINS, dtype=torch.float32)
        C.memmove(chk.data_ptr(), h_out.value + f0 * BINS * 4, nfr * BINS * 4)
        same = same and bool(torch.equal(chk.view(nfr, BINS), out[f0:f0 + nfr].cpu()))
    lib.kopek_host_free(h_in)
    lib.kopek_host_free(h_out)
    ceiling = copy_ceiling(c, min(in_bytes, 1 << 30), min(out_bytes, (1 << 30) * BINS // N_FFT))

    line = None
    if rank == 0:
        kernel_ms = statistics.mean(per_launch_ms)
        achieved = frames * ALGO_BYTES_PER_FRAME / (kernel_ms * 1e-3) / 1e9
        traffic, traffic_note = _ncu_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": max_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu": frames, "n": N_FFT, "hop": N_FFT, "window": "hann",
                       "output": "mag", "l2": "inputs (4 GiB/GPU) and outputs (2.1 GB/GPU) exceed the 126 MB L2",
                       "parallelism": f"frame-sharded x{world}, no data-path collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": c.peak, "unit": "GB/s",
                         "frac": achieved / c.peak, "traffic": traffic, "traffic_source": traffic_note,
                         "peak_source": peak_src,
                         "kernel": "k1w_1024b_kernel<MAG,WIN> (warp per frame, two shared-memory exchanges, LDG register prefetch)",
                         "kernel_src_sha": headline_kernel_fingerprint(), "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": frames * ALGO_BYTES_PER_FRAME,
                         "compute_roofs": compute_roofs("k1w_1024b<MAG,WIN>", frames, kernel_ms,
                                                        float(clocks.get("sm_max_mhz") or 1965.0)),
                         "sustained": None if sustained is None else dict(
                             sustained, frac=frames * ALGO_BYTES_PER_FRAME / (sustained["ms_per_step"] * 1e-3) / 1e9 / c.peak)},
            "parity": parity,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                    "frames_per_step": ef, "steps": e2e_steps, "launches_per_step": e2e_launches,
                    "matches_device_result": same,
                    "api": "kopek_stft_exec_host (pinned host buffers, chunked H2D/kernel/D2H on 3 streams)",
                    "copy_ceiling": ceiling, "frac_of_copy_ceiling": e2e_value / ceiling,
                    "copy_ceiling_how": "every rank at once, no kernel: the step's input H2D and its output D2H from pinned "
                                        "buffers on two streams, as one copy per direction and as 16 MiB pieces -- the "
                                        "better of the two (max over ranks)",
                    "cpus_near_gpu": near},
            "gpu_launches": launches,
            "clocks": clocks,
        }
    del out, x
    torch.cuda.empty_cache()

    # =============================================================== every BASELINE.json config ================
    configs = {}
    sm_max = float(clocks.get("sm_max_mhz") or 1965.0)
    if not args.no_configs:
        if rank == 0:
            import oracle
            oracle.build()
            configs["cfg0"] = bench_config_cfg0(c)
        barrier()
        # cfg1: 60 s of 48 kHz chirp, Hann 2048 / hop 512, dB -- one GPU's job: every rank runs it, max reported
        chirp = torch.from_numpy(signals.chirp(60 * 48000, 48000.0, 20.0, 20000.0)).to(device)
        configs["cfg1"] = bench_stft_config(
            c, "cfg1", 2048, 512, WINDOW_HANN, OUT_DB, chirp, 5622, "k1m_kernel<4> (64 lanes per frame)",
            "cfg1: 60 s x 48 kHz chirp 20 Hz -> 20 kHz, Hann 2048 / hop 512, dB, 1 GPU (5622 frames)",
            check_frames=5622, reps=200, copies=8)
        configs["cfg1"]["note"] = "34.6 MB per launch: launch/latency bound, the roofline fraction is not the figure of merit"
        configs["cfg1"]["compute_roofs"] = compute_roofs("k1m<4,DB,cos-sum>", 5622, configs["cfg1"]["ms"], sm_max)
        del chirp
        # cfg2: 1 000 000 x 256-point white noise |X|, STRONG-scaled: rank r owns frames [r F / R, (r + 1) F / R)
        F2 = 1_000_000
        f0, f1 = (rank * F2) // world, ((rank + 1) * F2) // world
        noise = synth(f1 - f0, 256, f0, wave=fft.WAVE_SILENCE, noise=fft.NOISE_WHITE, noise_gain=1.0, seed=42)
        r2 = bench_stft_config(c, "cfg2", 256, 256, WINDOW_RECT, OUT_MAG, noise, f1 - f0,
                               "k1w_256_kernel (8 lanes per frame)",
                               f"cfg2: 1M independent 256-pt FFTs of N(0,1) noise, |X|, strong-scaled over {world} GPU(s)",
                               check_frames=1024, reps=20 if world < 4 else 50, copies=max(1, min(8, world)))
        r2["scaling"] = "strong"
        r2["job_algorithmic_bytes"] = F2 * (256 * 4 + 129 * 4)
        r2["job_samples_per_s"] = F2 * 256 / (r2["ms"] * 1e-3)
        r2["frac_per_gpu"] = r2["frac"]
        configs["cfg2"] = r2
        del noise
        # the modes the reference's callers actually iterate (examples/graph/utils.rs:47-63 walks ALL N bins) and the
        # unaligned case, on the headline frame size: every rank runs them on 2^19 frames, max over ranks reported
        xm = synth(1 << 19, N_FFT, (1 << 24) + rank * (1 << 19), **m_kw)
        configs["full_rows_1024"] = bench_stft_config(
            c, "full_rows_1024", N_FFT, N_FFT, WINDOW_RECT, OUT_MAG, xm, 1 << 19,
            "k1w_1024b_kernel<MAG,rect,FULL> (mirror stored beside the computed half)",
            "FULL rows: 2^19 frames x 1024-pt rect |X|, all 1024 bins per row (what graph/pong iterate)", check_frames=128,
            full=True)
        hop_odd = N_FFT - 1
        fr_odd = 1 + (xm.numel() - N_FFT) // hop_odd
        configs["odd_hop_1024"] = bench_stft_config(
            c, "odd_hop_1024", N_FFT, hop_odd, WINDOW_HANN, OUT_MAG, xm, fr_odd,
            "k1w_1024b_kernel<MAG,WIN,VEC=false> (32-bit sample loads)",
            f"odd hop: {fr_odd} frames x 1024-pt Hann |X|, hop 1023 (frames start at any 4-byte aligned sample)",
            check_frames=128)
        del xm
        # cfg3: one 2^24-point c2c f32 transform (four-step), rank 0's GPU; cuFFT on the same buffer beside it
        configs["cfg3"] = bench_cfg3(c) if rank == 0 else None
        barrier()
        # cfg4: 8 h x 48 kHz per channel, Hann 4096 / hop 1024, |X| -- one FULL-SIZE channel per GPU (8 GPUs = configs[4])
        ns4 = 8 * 3600 * 48000
        frames4 = 1 + (ns4 - 4096) // 1024
        x4 = synth(ns4 // 4096, 4096, rank * (ns4 // 4096), wave=fft.WAVE_SINE, freq0=55.0, freq_step=7.0, freq_mod=3000,
                   noise=fft.NOISE_WHITE, noise_gain=0.05, seed=7 + rank)
        r4 = bench_stft_config(c, "cfg4", 4096, 1024, WINDOW_HANN, OUT_MAG, x4, frames4,
                               "k1m_kernel<8> (128 lanes per frame)",
                               f"cfg4: 8 h x 48 kHz x {world} channel(s) (one per GPU), Hann 4096 / hop 1024, |X|: "
                               f"{frames4} frames x 2049 bins per channel", check_frames=64, reps=5)
        r4["compute_roofs"] = compute_roofs("k1m<8,MAG,cos-sum>", frames4, r4["ms"], sm_max)
        r4["channels"] = world
        r4["job_transformed_samples_per_s"] = r4["transformed_samples_per_s"] * world
        r4["fp32_note"] = ("hop = N/4: every sample is transformed 4 times, so this shape is bound by FP32 issue, not by "
                           "HBM; transformed samples/s is the figure of merit")
        configs["cfg4"] = r4
        if world > 1:
            r4["gather_to_rank0"] = bench_cfg4_gather(c, x4, frames4)
        del x4
        torch.cuda.empty_cache()

    # =============================================================== multi-GPU legs ============================
    gather = dist_leg = mgpu = None
    if world > 1 and not args.no_multi:
        gather = bench_gather(c, m_kw)
        if world in (2, 4, 8):
            dist_leg = bench_dist_fft(c)
        torch.cuda.empty_cache()
        barrier()
        # the C ABI's single-process multi-GPU path: rank 0 drives ALL devices, the other ranks wait on the host
        if rank == 0:
            try:
                mgpu = bench_mgpu(c, m_kw)
            except Exception as e:          # report, do not lose the line
                mgpu = {"error": repr(e)}
        dist.barrier(group=cpu_group)

    if rank == 0:
        line["configs"] = configs
        if gather is not None:
            line["gather"] = gather
        if dist_leg is not None:
            line["dist_fft"] = dist_leg
        if mgpu is not None:
            line["mgpu"] = mgpu
        if world == 1 and not args.no_cpu:
            os.sched_setaffinity(0, all_cpus)            # the CPU port gets every host core
            line["cpu_baseline"] = cpu_baseline()
            line["cpu_baseline_1core"] = cpu_baseline(target_s=6.0, cores=1)
        print(json.dumps(line), file=out_stream, flush=True)
    plan.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def bench_cfg3(c):
    """configs[3]: single 2^24-point c2c f32 transform.  Parity: against the f64 oracle at 2^16 points (what the oracle
    finishes in seconds) and against an f64 library transform of the whole 2^24-point signal (second opinion)."""
    import numpy as np

    import oracle
    torch = c.torch
    from kopek_b200 import fft
    stream = torch.cuda.current_stream().cuda_stream
    res = {"workload": "cfg3: single 2^24-point complex f32 FFT (two-tone + N(0,1) noise, seed 7), four-step, 1 GPU",
           "kernel": "k3_pass1<4096> + k3_pass2_r64 (TMA-staged tiles, two HBM passes)"}
    g = torch.Generator(device="cpu").manual_seed(7)
    sizes = {}
    for lg in (20, 22, 23, 24):
        n = 1 << lg
        t = torch.arange(n, dtype=torch.float64)
        sig = (torch.sin(2 * math.pi * 0.01 * t) + 0.5 * torch.sin(2 * math.pi * 0.2503 * t)).to(torch.float32)
        a = torch.view_as_complex(torch.stack([sig + torch.randn(n, generator=g), torch.randn(n, generator=g)], dim=1)
                                  .contiguous()).to(c.device)
        o = torch.empty_like(a)
        ms = time_launches(torch, lambda i: fft.fft_large_device(a.data_ptr(), o.data_ptr(), n, c.local, stream), 20)
        lib_ms = time_launches(torch, lambda i: torch.fft.fft(a), 20)
        g_ms = g_lib = None
        if c.world == 1:            # (a capture beside live NCCL communicators is not worth the risk to the line)
            g_ms = time_graph_replays(torch, lambda st: fft.fft_large_device(a.data_ptr(), o.data_ptr(), n, c.local, st))
            g_lib = time_graph_replays(torch, lambda st: torch.fft.fft(a))
        want = torch.fft.fft(a.to(torch.complex128))
        err = float(((o.to(torch.complex128) - want).abs().max() / want.abs().max()).item())
        sizes[f"2^{lg}"] = {"ms": ms, "cufft_c2c_f32_ms": lib_ms, "speed_vs_cufft": lib_ms / ms,
                            "max_rel_err_vs_f64_library": err,
                            "graph_replay_ms": g_ms, "cufft_graph_replay_ms": g_lib,
                            "achieved_gbs": 2 * n * 8 / (ms * 1e-3) / 1e9, "frac": 2 * n * 8 / (ms * 1e-3) / 1e9 / c.peak}
        del a, o, want
    res["sizes"] = sizes
    res.update({k: sizes["2^24"][k] for k in ("ms", "cufft_c2c_f32_ms", "achieved_gbs", "frac", "max_rel_err_vs_f64_library")})
    res["algorithmic_bytes"] = 2 * (1 << 24) * 8
    res["two_pass_ceiling_frac"] = 0.5
    res["l2"] = "input + workspace + output = 403 MB > 126 MB L2"
    res["cufft_note"] = "torch.fft.fft on the same complex64 buffer (cuFFT C2C f32): the library bar, never the product"
    res["timing"] = ("ms / cufft_c2c_f32_ms: 20 eager back-to-back calls per size, CUDA events; graph_replay_ms: the same "
                     "calls captured into one CUDA graph and replayed (no host work between launches; N = 1 only)")
    # the oracle itself at a size it finishes in seconds: 2^16 points through the same four-step kernels
    n = 1 << 16
    rng = np.random.default_rng(7)
    z = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    z[5] += 300.0
    a = torch.from_numpy(z).to(c.device)
    o = torch.empty_like(a)
    fft.fft_large_device(a.data_ptr(), o.data_ptr(), n, c.local, stream)
    torch.cuda.synchronize()
    got = o.cpu().numpy().astype(np.complex128)
    tol = 1e-5 * math.log2(n)
    par = {"n": n, "tolerance": tol}
    for prec, dt in ((64, np.complex128), (32, np.complex64)):
        ref = oracle.fft(z.astype(dt), dtype=dt).astype(np.complex128)
        par[f"max_rel_err_vs_f{prec}_oracle"] = float(np.abs(got - ref).max() / np.abs(ref).max())
        if prec == 64:
            par["peak_bin_mismatches"] = int(np.abs(got).argmax() != np.abs(ref).argmax())
    par["ok"] = bool(par["max_rel_err_vs_f64_oracle"] <= tol and par["max_rel_err_vs_f32_oracle"] <= tol
                     and par["peak_bin_mismatches"] == 0)
    res["parity"] = par
    return res


def bench_cfg4_gather(c, x4, frames4):
    """configs[4] verbatim: the whole spectrogram (world x 11 GB) gathered to rank 0 with NCCL; per-block checksums taken
    before the transfer are compared with the gathered rows."""
    torch, dist = c.torch, c.dist
    from kopek_b200 import OUT_MAG, WINDOW_HANN, StftPlan
    from kopek_b200.sharded import gather_spectra
    stream = torch.cuda.current_stream().cuda_stream
    bins = 2049
    with StftPlan(4096, 1024, WINDOW_HANN, OUT_MAG, device=c.local) as plan:
        local = torch.empty((frames4, bins), device=c.device, dtype=torch.float32)
        plan.exec_ptr(x4.data_ptr(), x4.numel(), local.data_ptr(), stream)
        torch.cuda.synchronize()
    total = frames4 * c.world
    # rank 0 needs the whole spectrogram beside its own shard: decide together whether it fits (never die of OOM here)
    torch.cuda.empty_cache()
    free = torch.cuda.mem_get_info()[0]
    fits = [bool(free > total * bins * 4 + (6 << 30))]
    dist.broadcast_object_list(fits, src=0)
    if not fits[0]:
        del local
        return {"skipped": f"rank 0 has {free / 1e9:.0f} GB free, the gathered spectrogram needs {total * bins * 4 / 1e9:.0f} GB"}
    sums = block_checksums(torch, local)
    all_sums = [None] * c.world
    dist.all_gather_object(all_sums, sums)
    # uneven frame_range would split differently: every rank holds exactly frames4 rows, so ranges are r * frames4
    full = None
    times = []
    for it in range(3):
        full = None                       # hand the previous 88 GB back before the next one is allocated
        c.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        got = gather_spectra(local, total, dst=0)
        e1.record()
        torch.cuda.synchronize()
        times.append(max_over_ranks(c, e0.elapsed_time(e1)))
        if c.rank == 0:
            full = got
        del got
    res = {"what": f"{c.world} x {frames4} x {bins} f32 = {total * bins * 4 / 1e9:.1f} GB into rank 0 (grouped NCCL send/recv)",
           "first_call_ms": times[0], "ms": min(times[1:]),
           "gbs_into_rank0": (c.world - 1) * frames4 * bins * 4 / (min(times[1:]) * 1e-3) / 1e9}
    if c.rank == 0:
        okay = all(block_checksums(torch, full[r * frames4:(r + 1) * frames4]) == tuple(all_sums[r])
                   for r in range(c.world))
        res["checksums_match_source_shards"] = bool(okay)
    del full, local
    torch.cuda.empty_cache()
    return res


def bench_gather(c, m_kw):
    """ONE spectrogram of the headline shape on rank 0 (2^18 frames per rank, rows padded to 516 floats so that the
    bulk-store epilogue applies): (a) transform, then grouped NCCL send/recv; (b) every rank's kernel stores straight
    into rank 0's buffer over NVLink (one 2048-byte bulk store per frame) -- transform and transfer are one kernel.
    Both are compared bit for bit with rank 0 transforming all frames alone."""
    torch, dist = c.torch, c.dist
    from kopek_b200 import OUT_MAG, WINDOW_HANN, StftPlan, fft
    from kopek_b200.sharded import PeerSpectrogram, gather_spectra
    stream = torch.cuda.current_stream().cuda_stream
    fr, pitch = 1 << 18, 516
    total = fr * c.world

    def synth(count, first):
        x = torch.empty(count * N_FFT, device=c.device, dtype=torch.float32)
        fft.synth_frames_device(x.data_ptr(), count, N_FFT, 48000.0, first_frame=first, device=c.local, stream=stream, **m_kw)
        torch.cuda.synchronize()
        return x

    x = synth(fr, (1 << 22) + c.rank * fr)
    res = {"what": f"{c.world} x 2^18 frames x 1024-pt Hann |X|, rows of 516 f32, gathered to rank 0",
           "bytes_into_rank0": (c.world - 1) * fr * pitch * 4}
    with StftPlan(N_FFT, N_FFT, WINDOW_HANN, OUT_MAG, out_pitch=pitch, device=c.local) as plan, \
            StftPlan(N_FFT, N_FFT, WINDOW_HANN, OUT_MAG, out_pitch=pitch, device=c.local, bulk_store=True) as bplan:
        local = torch.empty((fr, pitch), device=c.device, dtype=torch.float32)
        ref = None
        if c.rank == 0:                     # the 1-GPU result of the whole job
            xa = synth(total, 1 << 22)
            ref = torch.empty((total, pitch), device=c.device, dtype=torch.float32)
            plan.exec_ptr(xa.data_ptr(), xa.numel(), ref.data_ptr(), stream)
            torch.cuda.synchronize()
            del xa
        # (0) transform alone
        res["transform_only_ms"] = max_over_ranks(c, time_launches(
            torch, lambda i: plan.exec_ptr(x.data_ptr(), x.numel(), local.data_ptr(), stream), 10))
        # (a) transform + NCCL gather
        full, t = None, []
        for it in range(4):
            full = None
            c.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.exec_ptr(x.data_ptr(), x.numel(), local.data_ptr(), stream)
            got = gather_spectra(local, total, dst=0)
            e1.record()
            torch.cuda.synchronize()
            t.append(max_over_ranks(c, e0.elapsed_time(e1)))
            if c.rank == 0:
                full = got
            del got
        res["nccl"] = {"ms": min(t[1:]), "first_call_ms": t[0],
                       "gbs_into_rank0": res["bytes_into_rank0"] / (min(t[1:]) * 1e-3) / 1e9}
        if c.rank == 0:
            res["nccl"]["bit_identical_to_1gpu"] = bool(torch.equal(full[:, :BINS], ref[:, :BINS]))
        del full
        # (b) fused: peer-direct bulk stores into rank 0's buffer
        peer = PeerSpectrogram(total, pitch, 4, c.local, dst=0)
        t = []
        for it in range(4):
            c.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            bplan.exec_ptr(x.data_ptr(), x.numel(), peer.row_ptr(c.rank * fr), stream)
            e1.record()
            peer.finish()
            t.append(max_over_ranks(c, e0.elapsed_time(e1)))
        res["peer_direct_bulk_store"] = {"ms": min(t[1:]), "gbs_into_rank0": res["bytes_into_rank0"] / (min(t[1:]) * 1e-3) / 1e9}
        if c.rank == 0:
            got = torch.empty((total, pitch), device=c.device, dtype=torch.float32)
            peer.copy_to(got.data_ptr())
            res["peer_direct_bulk_store"]["bit_identical_to_1gpu"] = bool(torch.equal(got[:, :BINS], ref[:, :BINS]))
            del got
        dist.barrier()
        if c.rank != 0:
            peer.close()
        dist.barrier()
        if c.rank == 0:
            peer.close()
        del ref, local
    torch.cuda.empty_cache()
    return res


def bench_dist_fft(c):
    """One complex f32 transform of world * 2^22 points over all GPUs (local four-step + ONE exchange kernel that reads
    the peers' partial spectra over NVLink), against an f64 transform of the whole signal."""
    torch, dist = c.torch, c.dist
    from kopek_b200.sharded import DistributedFFT, dist_fft_layout
    n = c.world << 22
    g = torch.Generator().manual_seed(22)
    x = torch.view_as_complex(torch.randn(n, 2, generator=g))
    x[5 * c.world + 1] += 300.0
    src, dst = dist_fft_layout(n, c.world, c.rank)
    mine = x[torch.from_numpy(src)].to(c.device).contiguous()
    out = torch.empty((c.world, n // c.world // c.world), dtype=torch.complex64, device=c.device)
    stream = torch.cuda.current_stream().cuda_stream
    f = DistributedFFT(n, c.local)
    f.forward(mine.data_ptr(), out.data_ptr(), stream)
    want = torch.fft.fft(x.to(c.device).to(torch.complex128))
    mine_want = want[torch.from_numpy(dst.reshape(-1)).to(c.device)].reshape(dst.shape)
    err = float(((out.to(torch.complex128) - mine_want).abs().max() / want.abs().max()).item())
    err = max_over_ranks(c, err)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        f.forward(mine.data_ptr(), out.data_ptr(), stream)
    ms_call = max_over_ranks(c, (time.perf_counter() - t0) / reps * 1e3)
    c.barrier()
    ms_x = max_over_ranks(c, time_launches(torch, lambda i: f.combine_only(out.data_ptr(), stream), 10))
    c.barrier()
    f.close()
    tol = 1e-5 * math.log2(n)
    return {"n": n, "ranks": c.world, "max_rel_err_vs_f64": err, "tolerance": tol, "ok": bool(err <= tol),
            "ms_per_call_with_barriers": ms_call, "exchange_kernel_ms": ms_x,
            "nvlink_read_gbs_per_gpu": (c.world - 1) / c.world * (n // c.world) * 8 / (ms_x * 1e-3) / 1e9}


def bench_mgpu(c, m_kw):
    """kopek_mgpu_* from ONE process (rank 0) over all GPUs of the job: host buffer in / spectrogram out, device
    shards + fused gather; every result compared bit for bit with the one-device plan."""
    import ctypes as C

    torch = c.torch
    from kopek_b200 import OUT_MAG, WINDOW_HANN, StftPlan, _lib, fft
    from kopek_b200.sharded import MultiGpuStft
    lib = _lib.load()
    devs = list(range(c.world))
    fr_per, pitch = 1 << 17, 516
    total = fr_per * c.world
    stream = torch.cuda.current_stream().cuda_stream
    x = torch.empty(total * N_FFT, device=c.device, dtype=torch.float32)
    fft.synth_frames_device(x.data_ptr(), total, N_FFT, 48000.0, first_frame=1 << 23, device=c.local, stream=stream, **m_kw)
    torch.cuda.synchronize()
    res = {"devices": devs, "what": f"rank 0 alone drives {c.world} GPUs through the C ABI: {total} frames x 1024-pt Hann |X|"}
    with StftPlan(N_FFT, N_FFT, WINDOW_HANN, OUT_MAG, out_pitch=pitch, device=c.local) as plan:
        ref = torch.empty((total, pitch), device=c.device, dtype=torch.float32)
        plan.exec_ptr(x.data_ptr(), x.numel(), ref.data_ptr(), stream)
        torch.cuda.synchronize()
    with MultiGpuStft(devs, N_FFT, N_FFT, WINDOW_HANN, OUT_MAG, out_pitch=pitch) as m:
        # host buffer in, host spectrogram out
        h_in, h_out = C.c_void_p(), C.c_void_p()
        _lib.check(lib.kopek_host_alloc(C.byref(h_in), total * N_FFT * 4))
        _lib.check(lib.kopek_host_alloc(C.byref(h_out), total * pitch * 4))
        xh = x.cpu()
        C.memmove(h_in.value, xh.data_ptr(), total * N_FFT * 4)
        del xh
        m.exec_host_ptr(h_in.value, total * N_FFT, h_out.value)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            m.exec_host_ptr(h_in.value, total * N_FFT, h_out.value)
        dt = (time.perf_counter() - t0) / reps
        got = torch.empty((total, pitch), dtype=torch.float32)
        C.memmove(got.data_ptr(), h_out.value, total * pitch * 4)
        res["exec_host"] = {"samples_per_s": total * N_FFT / dt, "ms": dt * 1e3,
                            "bit_identical_to_1gpu": bool(torch.equal(got[:, :BINS], ref[:, :BINS].cpu()))}
        lib.kopek_host_free(h_in)
        lib.kopek_host_free(h_out)
        del got
        # device-resident shards: transform alone, fused gather, copy-engine gather, NCCL gather
        ins, outs, counts, nfs = [], [], [], []
        for r, d in enumerate(devs):
            s0, cnt, f0, nf = m.shard(total * N_FFT, r)
            ins.append(x[s0:s0 + cnt].to(f"cuda:{d}"))
            outs.append(torch.empty((nf, pitch), device=f"cuda:{d}", dtype=torch.float32))
            counts.append(cnt)
            nfs.append(nf)
        full = torch.empty((total, pitch), device=c.device, dtype=torch.float32)
        for d in devs:
            torch.cuda.synchronize(d)
        ip, op = [t.data_ptr() for t in ins], [t.data_ptr() for t in outs]

        def timed(fn, reps=5):
            fn()
            m.sync()
            best = 1e30
            for _ in range(reps):
                fn()
                best = min(best, m.sync())
            return best

        res["exec_ms"] = timed(lambda: m.exec_ptrs(ip, counts, op))
        res["samples_per_s_device_resident"] = total * N_FFT / (res["exec_ms"] * 1e-3)
        bytes_in = (c.world - 1) * fr_per * pitch * 4
        for name, fn in (("exec_gathered", lambda: m.exec_gathered_ptr(ip, counts, full.data_ptr())),
                         ("exec_then_gather_copy", lambda: (m.exec_ptrs(ip, counts, op), m.gather_ptrs(op, nfs, full.data_ptr(), _lib.GATHER_COPY))),
                         ("exec_then_gather_nccl", lambda: (m.exec_ptrs(ip, counts, op), m.gather_ptrs(op, nfs, full.data_ptr(), _lib.GATHER_NCCL)))):
            full.fill_(float("nan"))
            torch.cuda.synchronize(c.local)
            try:
                ms = timed(fn)
            except _lib.KopekError as e:
                res[name] = {"error": str(e)}
                continue
            res[name] = {"ms": ms, "gbs_into_device0": bytes_in / (ms * 1e-3) / 1e9,
                         "bit_identical_to_1gpu": bool(torch.equal(full[:, :BINS], ref[:, :BINS]))}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--sustained-steps", type=int, default=300,
                    help="extra back-to-back launches after the timed region, reported as roofline.sustained (0 = skip)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES_PER_GPU, help="frames per GPU (default: the headline 2^20)")
    ap.add_argument("--e2e-frames", type=int, default=E2E_FRAMES)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline legs")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config legs (cfg0..cfg4)")
    ap.add_argument("--no-multi", action="store_true", help="skip the gather / dist_fft / mgpu legs at N > 1")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the headline output")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
