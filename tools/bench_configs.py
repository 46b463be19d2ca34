#!/usr/bin/env python3
"""Time every BASELINE.json config (not only the headline) on one GPU and print a markdown table.

  configs[0] single 1024-pt fft() drop-in call (latency, host buffers)         -> K2
  configs[1] 60 s 48 kHz chirp, Hann 2048, hop 512, dB                         -> K1
  configs[2] 1M x 256-pt white noise, |X|                                       -> K1
  configs[3] single 2^24-point c2c f32                                          -> K3
  configs[4] (per-GPU shard) 8ch x 8h x 48 kHz, Hann 4096, hop 1024, |X|       -> K1 (1/64 of the job here)
  M          2^20 x 1024 Hann |X|                                               -> K1W
Roofline denominator: MEASURED_PEAKS.json hbm_gbs.  Algorithmic bytes as in SURVEY.md 8d.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from kopek_b200 import OUT_DB, OUT_MAG, WINDOW_HANN, WINDOW_RECT, StftPlan, fft, signals  # noqa: E402

try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
st = torch.cuda.current_stream().cuda_stream


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


rows = []


def stft_case(name, n, hop, window, output, n_samples, gen=None, bins=0, offset=0):
    """offset: samples skipped at the start of the buffer (1 = a base pointer that is only 4-byte aligned)."""
    x = torch.randn(n_samples + offset, device="cuda") if gen is None else torch.from_numpy(gen).cuda()
    plan = StftPlan(n, hop, window, output, bins=bins)
    frames = plan.frames(n_samples)
    out = torch.empty((frames, plan.pitch), device="cuda")
    time.sleep(1.0)                      # every row starts from a cooled board: burst figures, comparable between rows
    ms = timed(lambda: plan.exec_ptr(x.data_ptr() + 4 * offset, n_samples, out.data_ptr(), st))
    by = ((frames - 1) * hop + n) * 4 + frames * plan.bins * 4
    rows.append((name, f"{frames} frames", ms, by, frames * n / ms / 1e6))
    plan.close()
    del x, out


# M
stft_case("M: 2^20 x 1024 Hann |X| (K1W-b)", 1024, 1024, WINDOW_HANN, OUT_MAG, (1 << 20) * 1024)
# configs[1]
stft_case("cfg1: 60 s chirp, Hann 2048 hop 512 dB (K1M)", 2048, 512, WINDOW_HANN, OUT_DB, 2880000,
          gen=signals.chirp(2880000, 48000.0, 20.0, 20000.0))
# configs[2]
stft_case("cfg2: 1M x 256 white noise |X| (K1W-256)", 256, 256, WINDOW_RECT, OUT_MAG, 1000000 * 256)
# configs[4], one GPU's shard of one channel-hour-ish: 2^27 samples
stft_case("cfg4 shard: 2^27 samples, Hann 4096 hop 1024 |X| (K1M)", 4096, 1024, WINDOW_HANN, OUT_MAG, 1 << 27)
stft_case("512-pt: 2^21 x 512 Hann |X| (K1W-512)", 512, 512, WINDOW_HANN, OUT_MAG, (1 << 21) * 512)
stft_case("2048-pt: 2^19 x 2048 Hann |X| (K1M)", 2048, 2048, WINDOW_HANN, OUT_MAG, (1 << 19) * 2048)
stft_case("4096-pt: 2^18 x 4096 Hann |X| (K1M)", 4096, 4096, WINDOW_HANN, OUT_MAG, (1 << 18) * 4096)
stft_case("1024-pt STFT hop 256: 2^28 samples Hann |X| (K1W-b)", 1024, 256, WINDOW_HANN, OUT_MAG, 1 << 28)
# what the reference's callers iterate: ALL N bins per row (examples/graph/utils.rs:47-63) -- round 2: on the lane-group kernels
stft_case("FULL rows 1024: 2^20 x 1024 rect |X|, 1024 bins/row (K1W-b)", 1024, 1024, WINDOW_RECT, OUT_MAG, (1 << 20) * 1024, bins=1)
stft_case("FULL rows 256: 2^22 x 256 rect |X| (K1W-256)", 256, 256, WINDOW_RECT, OUT_MAG, (1 << 22) * 256, bins=1)
stft_case("FULL rows 512: 2^21 x 512 Hann |X| (K1W-512)", 512, 512, WINDOW_HANN, OUT_MAG, (1 << 21) * 512, bins=1)
stft_case("FULL rows 2048: 2^19 x 2048 Hann |X| (K1M)", 2048, 2048, WINDOW_HANN, OUT_MAG, (1 << 19) * 2048, bins=1)
stft_case("FULL rows 4096: 2^18 x 4096 Hann |X| (K1M)", 4096, 4096, WINDOW_HANN, OUT_MAG, (1 << 18) * 4096, bins=1)
# odd hops / a base pointer that is only 4-byte aligned (32-bit loads on the same kernels)
stft_case("odd hop 1023, 1024-pt Hann |X| (K1W-b, 32-bit loads)", 1024, 1023, WINDOW_HANN, OUT_MAG, (1 << 20) * 1023 + 1)
stft_case("unaligned base (+1 sample), 1024-pt Hann |X| (K1W-b)", 1024, 1024, WINDOW_HANN, OUT_MAG, (1 << 20) * 1024, offset=1)
stft_case("odd hop 2047, 2048-pt Hann |X| (K1M, 32-bit loads)", 2048, 2047, WINDOW_HANN, OUT_MAG, (1 << 19) * 2047 + 1)
stft_case("odd hop 4095, 4096-pt Hann |X| (K1M, 32-bit loads)", 4096, 4095, WINDOW_HANN, OUT_MAG, (1 << 18) * 4095 + 1)
stft_case("odd hop 255, 256-pt rect |X| (K1W-256, 32-bit loads)", 256, 255, WINDOW_RECT, OUT_MAG, (1 << 22) * 255 + 1)
# fused band epilogues (input bytes only)
stft_case("bands LOG: 2^20 x 1024 Hann -> 10 f32/frame (K1W-b)", 1024, 1024, WINDOW_HANN, 4, (1 << 20) * 1024)
stft_case("bands LOW: 2^20 x 1024 Hann -> 10 f32/frame (K1W-b)", 1024, 1024, WINDOW_HANN, 5, (1 << 20) * 1024)
# the library bar on the headline shape: cuFFT R2C through torch.fft.rfft, then a separate |.| kernel,
# window multiply included (three extra HBM round trips); reported beside the fused kernel, never part of the product
lib_rows = []
for n, frames in ((1024, 1 << 20), (256, 1000000), (4096, 1 << 18)):
    x = torch.randn(frames, n, device="cuda")
    w = torch.hann_window(n, periodic=True, device="cuda")
    ms = timed(lambda: torch.fft.rfft(x * w, dim=1).abs(), iters=10, warm=2)
    lib_rows.append((f"cuFFT R2C + window + abs (torch), {frames} x {n}", ms, frames * (n * 4 + (n // 2 + 1) * 4), frames * n / ms / 1e6))
    del x
torch.cuda.empty_cache()
# configs[3]
for lg in (20, 22, 24):
    n = 1 << lg
    a = torch.randn(n, dtype=torch.complex64, device="cuda")
    b = torch.empty_like(a)
    ms = timed(lambda: fft.fft_large_device(a.data_ptr(), b.data_ptr(), n, 0, st))
    rows.append((f"cfg3: single 2^{lg} c2c f32 four-step (K3)", "1 transform", ms, 2 * n * 8, n / ms / 1e6))
    ms = timed(lambda: torch.fft.fft(a))
    lib_rows.append((f"cuFFT C2C f32 (torch.fft.fft), single 2^{lg}", ms, 2 * n * 8, n / ms / 1e6))
# decoder-side callers: i16 stereo frames as kopek::decoder::decode returns them
dec_rows = []
nfr = 1 << 28
pcm = torch.randint(-32768, 32767, (nfr, 2), dtype=torch.int16, device="cuda")
o32 = torch.empty(nfr, device="cuda")
ms = timed(lambda: fft.pcm16_to_f32_device(pcm.data_ptr(), nfr, o32.data_ptr(), 2, 0, stream=st), iters=10)
dec_rows.append((f"pcm16 -> f32, channel 0 / 32767 (f64 division), 2^28 stereo frames", ms, nfr * 8))
ms = timed(lambda: fft.pcm16_scale_f32_device(pcm.data_ptr(), nfr, 0.00002, o32.data_ptr(), stream=st), iters=10)
dec_rows.append((f"pcm16 -> f32 track, 0.00002 * value (load_track_at_path), 2^28 values", ms, nfr * 6))
del o32
pl = StftPlan(1024, 1024, WINDOW_HANN, OUT_MAG)
fr = nfr // 1024
spec = torch.empty((fr, pl.pitch), device="cuda")
ms = timed(lambda: pl.exec_pcm16_ptr(pcm.data_ptr(), nfr, spec.data_ptr(), 0, st), iters=10)
dec_rows.append((f"pcm16 stereo -> Hann 1024 |X| fused (kopek_stft_exec_pcm16), 2^18 frames", ms, nfr * 4 + fr * 513 * 4))
pl.close()
del spec
hour = 3600 * 44100
ms = timed(lambda: fft.detect_bpm_device(pcm.data_ptr(), hour, 44100.0, stream=st), iters=5, warm=1)
dec_rows.append((f"detect_bpm, 1 h of 44.1 kHz stereo (sequential f32 folds, bit-exact)", ms, hour * 4))
ms = timed(lambda: fft.detect_bpm_device(pcm.data_ptr(), 124443, 44100.0, stream=st), iters=5, warm=1)
dec_rows.append((f"detect_bpm, 2.8 s clip (the reference's sample.wav size): latency", ms, 124443 * 4))
del pcm
torch.cuda.empty_cache()
# configs[0]: host-to-host latency of one drop-in call
x = (np.random.default_rng(0).standard_normal(1024)).astype(np.complex128)
fft.fft(x)
t0 = time.perf_counter()
for _ in range(200):
    fft.fft(x)
lat = (time.perf_counter() - t0) / 200 * 1e6

print(f"| config | units | ms | algorithmic GB/s | % of {PEAK:.1f} GB/s | G samples/s |\n|---|---|---:|---:|---:|---:|")
for name, units, ms, by, gs in rows:
    print(f"| {name} | {units} | {ms:.4f} | {by / ms / 1e6:.1f} | {100 * by / ms / 1e6 / PEAK:.1f} | {gs:.1f} |")
print("\n| library bar (not the product) | ms | algorithmic GB/s | % | G samples/s |\n|---|---:|---:|---:|---:|")
for name, ms, by, gs in lib_rows:
    print(f"| {name} | {ms:.4f} | {by / ms / 1e6:.1f} | {100 * by / ms / 1e6 / PEAK:.1f} | {gs:.1f} |")
print("\n| decoder-side callers | ms | algorithmic GB/s | % |\n|---|---:|---:|---:|")
for name, ms, by in dec_rows:
    print(f"| {name} | {ms:.4f} | {by / ms / 1e6:.1f} | {100 * by / ms / 1e6 / PEAK:.1f} |")
print(f"\ncfg0: one 1024-pt fft() drop-in call, host buffers in/out (K2): {lat:.1f} us per call")
