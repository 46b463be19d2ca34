#!/usr/bin/env python3
"""What the box's host<->device path delivers with NO kernel.

    python tools/pcie_ceiling.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_ceiling.py

Every rank, at the same time, moves the traffic of one e2e step of bench.py from / to pinned host buffers: H2D only,
D2H only, and both at once (4 bytes in + ~2 bytes out per sample, one cudaMemcpyAsync per direction per step on two
streams).  The "both" row is the ceiling `e2e.frac_of_copy_ceiling` in the bench line is quoted against; the other two
say which direction (or the host's memory system, when 8 ranks pull at once) is the limit.  One JSON line per case.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    in_bytes, out_bytes = 1 << 30, (1 << 30) * 513 // 1024
    h_in = torch.empty(in_bytes, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(out_bytes, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(in_bytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(out_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for case in ("h2d", "d2h", "both"):
        def once():
            if case in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if case in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
            s1.synchronize()
            s2.synchronize()
        once()
        barrier()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            once()
        barrier()
        t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item()) / reps
        if rank == 0:
            h2d = world * in_bytes / dt / 1e9 if case != "d2h" else 0.0
            d2h = world * out_bytes / dt / 1e9 if case != "h2d" else 0.0
            print(json.dumps({"case": case, "ranks": world, "ms": dt * 1e3, "h2d_gbs_total": h2d, "d2h_gbs_total": d2h,
                              "samples_per_s_equivalent": world * (in_bytes / 4) / dt if case != "d2h" else None,
                              "cpus": len(os.sched_getaffinity(0))}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
