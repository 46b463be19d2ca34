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

One JSON line on stdout (rank 0): value = whole-job samples/s with inputs
resident in HBM; e2e = the same metric through the host-buffer C-ABI call with
pinned host buffers (H2D + kernel + D2H inside the timed region); roofline of
the K1 kernel against MEASURED_PEAKS.json; cpu_baseline = the oracle port of
src/fft.rs timed on the host cores (rank 0, N = 1).
`--impl reference` times that CPU port alone with all host threads.
"""
from __future__ import annotations

import argparse
import json
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
E2E_FRAMES = 1 << 18
ALGO_BYTES_PER_FRAME = N_FFT * 4 + BINS * 4          # 6148: samples read once + magnitudes written once
METRIC = "complex samples/sec batched 1024-pt FFT+mag"
UNIT = "samples/s"
WORKLOAD = "M: 2^20 frames x 1024-pt real f32 per GPU, Hann window, |X| (513 f32/frame), non-overlapping"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _ncu_traffic():
    """dram bytes per K1 launch from the committed ncu capture, if there is one."""
    try:
        with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


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
    e2e leg are allocated, so that they come from the GPU's own NUMA node (eight ranks pulling 150 GB/s through one
    socket's memory controllers is what bounded the 8-GPU e2e figure).  Returns the number of CPUs, None if unknown."""
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


def cpu_port_rate(frames: int, threads: int):
    """Time the oracle port (restatement of src/fft.rs) on `frames` frames of the workload."""
    import oracle  # CPU baseline leg: the one place bench.py executes oracle/
    from kopek_b200 import signals
    x = signals.sine_frames(frames, N_FFT)
    t0 = time.perf_counter()
    oracle.stft_mag(x, N_FFT, N_FFT, oracle.WIN_HANN, output=oracle.OUT_MAG, threads=threads)
    dt = time.perf_counter() - t0
    return frames * N_FFT / dt, dt


def cpu_baseline(target_s: float = 18.0):
    import oracle
    oracle.build()
    cores = oracle.max_threads()
    probe = 256 * cores
    rate, _ = cpu_port_rate(probe, cores)
    frames = int(max(probe, min(rate * target_s / N_FFT, 4_000_000)))
    rate, dt = cpu_port_rate(frames, cores)
    return {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{frames} frames of the workload ({dt:.1f} s): oracle/ref_fft.c, f64 recursion of "
                      f"src/fft.rs + Hann + norm(), {cores} pthreads over independent frames"}


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


def make_input(torch, device, frames, first_frame, stream):
    """Workload M generated on the device by the library's own oscillator/noise sources (kopek_synth_frames):
    frame f = Oscillator::sine at 20*(1 + f mod 1000) Hz / 48 kHz from phase 0, + 0.1 * N(0,1), seed 1234."""
    from kopek_b200 import fft
    x = torch.empty(frames * N_FFT, device=device, dtype=torch.float32)
    fft.synth_frames_device(x.data_ptr(), frames, N_FFT, 48000.0, wave=fft.WAVE_SINE, freq0=20.0, freq_step=20.0,
                            freq_mod=1000, noise=fft.NOISE_WHITE, noise_gain=0.1, seed=1234, first_frame=first_frame,
                            device=device.index or 0, stream=stream)
    torch.cuda.synchronize()
    return x


def run_ours(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    from kopek_b200 import OUT_MAG, WINDOW_HANN, StftPlan, _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    out_stream = claim_stdout()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: kopek_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    all_cpus = os.sched_getaffinity(0)
    near = bind_near_gpu(_physical_gpu_index(local))
    frames = args.frames
    plan = StftPlan(N_FFT, N_FFT, WINDOW_HANN, OUT_MAG, device=local)
    stream = torch.cuda.current_stream().cuda_stream
    x = make_input(torch, device, frames, rank * frames, stream)
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
    t = torch.tensor([total_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    value = world * frames * N_FFT * args.steps / (max_ms * 1e-3)

    # ---- end to end: pinned host buffers -> kopek_stft_exec_host -> pinned host spectra
    lib = _lib.load()
    ef = min(args.e2e_frames, frames)
    h_in, h_out = C.c_void_p(), C.c_void_p()
    in_bytes, out_bytes = ef * N_FFT * 4, ef * BINS * 4
    _lib.check(lib.kopek_host_alloc(C.byref(h_in), in_bytes))
    _lib.check(lib.kopek_host_alloc(C.byref(h_out), out_bytes))
    x_host = x[:ef * N_FFT].cpu()
    C.memmove(h_in.value, x_host.data_ptr(), in_bytes)
    del x_host
    e2e_steps = max(3, min(args.steps, 10))
    plan.exec_host_ptr(h_in.value, ef * N_FFT, h_out.value)          # warm-up (allocates the chunk slots)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        plan.exec_host_ptr(h_in.value, ef * N_FFT, h_out.value)
    e2e_launches = plan.last_launches
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * ef * N_FFT * e2e_steps / float(te.item())
    # the spectra that came back over PCIe equal the device-resident ones
    chk = torch.empty(ef * BINS, dtype=torch.float32)
    C.memmove(chk.data_ptr(), h_out.value, out_bytes)
    same = bool(torch.equal(chk.view(ef, BINS), out[:ef].cpu()))
    lib.kopek_host_free(h_in)
    lib.kopek_host_free(h_out)

    if rank == 0:
        peak, peak_src = _peaks()
        kernel_ms = statistics.mean(per_launch_ms)
        achieved = frames * ALGO_BYTES_PER_FRAME / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": max_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu": frames, "n": N_FFT, "hop": N_FFT, "window": "hann",
                       "output": "mag", "l2": "inputs (4 GiB/GPU) and outputs (2.1 GB/GPU) exceed the 126 MB L2",
                       "parallelism": f"frame-sharded x{world}, no data-path collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": _ncu_traffic(), "peak_source": peak_src,
                         "kernel": "k1w_1024b_kernel<MAG,WIN> (warp per frame, two shared-memory exchanges, LDG register prefetch)", "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": frames * ALGO_BYTES_PER_FRAME,
                         "sustained": None if sustained is None else dict(
                             sustained, frac=frames * ALGO_BYTES_PER_FRAME / (sustained["ms_per_step"] * 1e-3) / 1e9 / peak)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                    "frames_per_step": ef, "steps": e2e_steps, "launches_per_step": e2e_launches,
                    "matches_device_result": same,
                    "api": "kopek_stft_exec_host (pinned host buffers, chunked H2D/kernel/D2H on 3 streams)"},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        line["e2e"]["cpus_near_gpu"] = near
        if world == 1 and not args.no_cpu:
            os.sched_setaffinity(0, all_cpus)            # the CPU port gets every host core
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), file=out_stream, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


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
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
