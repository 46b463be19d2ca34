#!/usr/bin/env python3
"""BASELINE.json configs[4] at FULL size under torchrun: an 8-hour 8-channel 48 kHz synthetic recording, STFT with a
Hann window of 4096 and hop 1024, sharded over the ranks (channel-major: 8 channels over R ranks), then the
spectrogram gathered to rank 0 with NCCL.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_cfg5.py [hours=8]

Transform time is CUDA-event time, max over ranks; the gather is reported separately (it is bounded by rank 0's
NVLink ingress).  The gathered rows are checked against per-rank checksums taken before the transfer.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from kopek_b200 import OUT_MAG, WINDOW_HANN, StftPlan  # noqa: E402
from kopek_b200.sharded import gather_spectra  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
hours = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
CH, N, HOP, SR = 8, 4096, 1024, 48000
per_ch = int(hours * 3600 * SR)
frames_ch = 1 + (per_ch - N) // HOP
my_ch = [c for c in range(CH) if c % world == rank]
st = torch.cuda.current_stream().cuda_stream
plan = StftPlan(N, HOP, WINDOW_HANN, OUT_MAG, device=local)
bins = plan.bins

# synthetic recording per channel, generated on the device: slow chirp + noise, seeded by the channel
xs = []
for c in my_ch:
    g = torch.Generator(device="cuda").manual_seed(1000 + c)
    x = torch.randn(per_ch, device="cuda", generator=g) * 0.1
    t = torch.arange(0, per_ch, device="cuda", dtype=torch.float32)
    x += torch.sin(t * (2 * 3.14159265 * (200.0 + 50.0 * c) / SR))
    del t
    xs.append(x)
outs = [torch.empty((frames_ch, bins), device="cuda") for _ in my_ch]


def transform():
    for x, o in zip(xs, outs):
        plan.exec_ptr(x.data_ptr(), per_ch, o.data_ptr(), st)


transform()
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 3
e0.record()
for _ in range(iters):
    transform()
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms_tr = t.item()

# checksums of the local blocks before they move
local_sum = torch.stack([o.double().sum() for o in outs]) if outs else torch.zeros(0, dtype=torch.float64, device="cuda")
sums = [torch.zeros(len([c for c in range(CH) if c % world == r]), dtype=torch.float64, device="cuda") for r in range(world)]
dist.all_gather(sums, local_sum) if all(s.numel() == local_sum.numel() for s in sums) else None

# gather: rank r's rows are its channels, concatenated; rank 0 ends with [CH * frames_ch, bins]
local_block = outs[0] if len(outs) == 1 else torch.cat(outs, 0)
total_rows = CH * frames_ch
ms_all = []
full = None
for it in range(3):          # the first call also pays NCCL's lazy peer connections and the 88 GB cudaMalloc on rank 0
    del full
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    full = gather_spectra(local_block, total_rows, dst=0)
    torch.cuda.synchronize(); dist.barrier()
    ms_all.append((time.perf_counter() - t0) * 1e3)
ms_ga = min(ms_all[1:])
ok = None
if rank == 0:
    ok = True
    rows_per_rank = total_rows // world
    for r in range(world):
        chs = [c for c in range(CH) if c % world == r]
        for i, c in enumerate(chs):
            got = full[r * rows_per_rank + i * frames_ch: r * rows_per_rank + (i + 1) * frames_ch].double().sum()
            ok = ok and bool(got == sums[r][i])
    samples_tr = CH * frames_ch * N
    algo = CH * (per_ch * 4 + frames_ch * bins * 4)
    print(json.dumps({
        "config": f"configs[4]: {hours} h x {CH} ch x {SR} Hz, Hann {N}, hop {HOP}, |X|", "ranks": world,
        "frames_total": CH * frames_ch, "bins": bins, "input_GB": CH * per_ch * 4 / 1e9,
        "spectrogram_GB": CH * frames_ch * bins * 4 / 1e9,
        "transform_ms_max_over_ranks": ms_tr, "transformed_T_samples_per_s": samples_tr / ms_tr / 1e9,
        "algorithmic_GBps_whole_job": algo / ms_tr / 1e6, "algorithmic_GBps_per_gpu": algo / ms_tr / 1e6 / world,
        "gather_ms_first_call": ms_all[0], "gather_ms": ms_ga, "gather_GBps_into_rank0": (world - 1) / world * CH * frames_ch * bins * 4 / ms_ga / 1e6,
        "gathered_rows_match_per_rank_checksums": ok}), flush=True)
dist.destroy_process_group()
