#!/usr/bin/env python3
"""torchrun tool: one spectrogram on rank 0 -- (a) transform then NCCL gather vs (b) peer-direct stores over NVLink.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_gather.py [frames_per_rank]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from kopek_b200 import OUT_MAG, WINDOW_HANN, StftPlan  # noqa: E402
from kopek_b200.sharded import PeerSpectrogram, gather_spectra  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
fpr = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
n = 1024
frames = fpr * world
x = torch.randn(fpr * n, device="cuda")
st = torch.cuda.current_stream().cuda_stream
# pitch 516: rows 16-byte aligned, so the peer path can use the bulk-store epilogue (one TMA store per frame)
plan = StftPlan(n, n, WINDOW_HANN, OUT_MAG, out_pitch=516, device=local)
plan_bulk = StftPlan(n, n, WINDOW_HANN, OUT_MAG, out_pitch=516, device=local, bulk_store=True)
local_out = torch.empty((fpr, plan.pitch), device="cuda")


def timed(fn, iters=5):
    fn()
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([(time.perf_counter() - t0) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item() * 1e3


def nccl_path():
    plan.exec_ptr(x.data_ptr(), x.numel(), local_out.data_ptr(), st)
    return gather_spectra(local_out, frames, dst=0)


peer = PeerSpectrogram(frames, plan.pitch, plan.elem_bytes, local, dst=0)


def peer_path():
    plan.exec_ptr(x.data_ptr(), x.numel(), peer.row_ptr(rank * fpr), st)
    torch.cuda.synchronize()


def peer_bulk_path():
    plan_bulk.exec_ptr(x.data_ptr(), x.numel(), peer.row_ptr(rank * fpr), st)
    torch.cuda.synchronize()


def local_only():
    plan.exec_ptr(x.data_ptr(), x.numel(), local_out.data_ptr(), st)


def local_bulk():
    plan_bulk.exec_ptr(x.data_ptr(), x.numel(), local_out.data_ptr(), st)


t_local, t_lbulk, t_nccl, t_peer, t_pbulk = timed(local_only), timed(local_bulk), timed(nccl_path), timed(peer_path), timed(peer_bulk_path)
full = nccl_path()
peer_bulk_path(); peer.finish()
if rank == 0:
    got = torch.empty((frames, plan.pitch), device="cuda")
    peer.copy_to(got.data_ptr())
    same = bool(torch.equal(got[:, :plan.bins], full[:, :plan.bins]))
    gb = (world - 1) * fpr * plan.bins * 4 / 1e9
    print(f"world={world} frames/rank={fpr}: transform only {t_local:.3f} ms (bulk store {t_lbulk:.3f}) | "
          f"transform + NCCL gather {t_nccl:.3f} ms ({gb / (t_nccl * 1e-3):.0f} GB/s) | peer-direct 32-bit stores "
          f"{t_peer:.3f} ms ({gb / (t_peer * 1e-3):.0f} GB/s) | peer-direct bulk stores {t_pbulk:.3f} ms "
          f"({gb / (t_pbulk * 1e-3):.0f} GB/s) | {gb:.2f} GB into rank 0 | identical={same}")
dist.barrier()
if rank != 0:
    peer.close()
dist.barrier()
if rank == 0:
    peer.close()
dist.destroy_process_group()
