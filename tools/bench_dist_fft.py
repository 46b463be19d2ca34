#!/usr/bin/env python3
"""torchrun tool: one transform of world * 2^lg points over the GPUs of one node (kopek_b200.sharded.DistributedFFT).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_dist_fft.py [lg=24]

Reports the local four-step transform, the exchange kernel (peer loads over NVLink + twiddle + radix-N butterfly) and
the whole call; CUDA-event times, max over ranks.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from kopek_b200 import fft  # noqa: E402
from kopek_b200.sharded import DistributedFFT  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
L = 1 << lg
n = world * L
x = torch.randn(L, dtype=torch.complex64, device="cuda")
tmp = torch.empty_like(x)
out = torch.empty((world, L // world), dtype=torch.complex64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
f = DistributedFFT(n, local)


def maxms(v):
    t = torch.tensor([v], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def ev_time(fn, iters):
    fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return maxms(e0.elapsed_time(e1) / iters)


f.forward(x.data_ptr(), out.data_ptr(), st)                                   # F_r in place everywhere
ms_local = ev_time(lambda: fft.fft_large_device(x.data_ptr(), tmp.data_ptr(), L, local, st), 10)
ms_comb = ev_time(lambda: f.combine_only(out.data_ptr(), st), 10)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(5):
    f.forward(x.data_ptr(), out.data_ptr(), st)
ms_call = maxms((time.perf_counter() - t0) / 5 * 1e3)
if rank == 0:
    pulled = (world - 1) / world * L * 8
    print(json.dumps({"ranks": world, "n": n, "log2_n": lg + world.bit_length() - 1, "local_transform_ms": ms_local,
                      "exchange_kernel_ms": ms_comb, "nvlink_read_GBps_per_gpu": pulled / ms_comb / 1e6,
                      "whole_call_ms_with_two_barriers": ms_call,
                      "algorithmic_GBps_whole_job_kernels_only": 2 * n * 8 / (ms_local + ms_comb) / 1e6}), flush=True)
f.close()
dist.destroy_process_group()
