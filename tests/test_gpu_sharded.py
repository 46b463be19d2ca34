"""GPU (>= 2 devices): frame-sharded STFT over one process per GPU + NCCL gather == the 1-GPU result, bit for bit."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from kopek_b200 import StftPlan, OUT_MAG, WINDOW_HANN, signals
from kopek_b200.sharded import sample_range, gather_spectra
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, hop = 1024, 256
x = signals.chirp(1_000_003, 48000.0, 20.0, 20000.0)            # same buffer on every rank
frames = 1 + (x.size - n) // hop
s0, s1, f0, f1 = sample_range(rank, world, x.size, n, hop)
st = torch.cuda.current_stream().cuda_stream
with StftPlan(n, hop, WINDOW_HANN, OUT_MAG, device=local) as plan:
    d = torch.from_numpy(x[s0:s1]).cuda()
    local_out = torch.empty((f1 - f0, plan.bins), device="cuda")
    assert plan.exec_ptr(d.data_ptr(), d.numel(), local_out.data_ptr(), st) == f1 - f0
    torch.cuda.synchronize()
    full = gather_spectra(local_out, frames, dst=0)
    if rank == 0:
        dx = torch.from_numpy(x).cuda()
        ref = torch.empty((frames, plan.bins), device="cuda")
        plan.exec_ptr(dx.data_ptr(), dx.numel(), ref.data_ptr(), st)
        torch.cuda.synchronize()
        assert full.shape == ref.shape and torch.equal(full, ref)      # shards are bit-identical to the 1-GPU run
        print("SHARDED_OK", world, frames)
    # peer-direct: every rank's kernel stores straight into rank 0's buffer over NVLink (no gather afterwards)
    from kopek_b200.sharded import PeerSpectrogram
    peer = PeerSpectrogram(frames, plan.pitch, plan.elem_bytes, local, dst=0)
    assert plan.exec_ptr(d.data_ptr(), d.numel(), peer.row_ptr(f0), st) == f1 - f0
    peer.finish()
    if rank == 0:
        got = torch.empty((frames, plan.pitch), device="cuda")
        peer.copy_to(got.data_ptr())
        assert torch.equal(got[:, :plan.bins], ref)
        print("PEER_DIRECT_OK")
    dist.barrier()
    if rank != 0:
        peer.close()
    dist.barrier()
    if rank == 0:
        peer.close()
# the same with 16-byte aligned rows and the bulk-store epilogue (one TMA store per frame crosses NVLink)
with StftPlan(n, hop, WINDOW_HANN, OUT_MAG, out_pitch=516, device=local, bulk_store=True) as bplan:
    peer = PeerSpectrogram(frames, 516, 4, local, dst=0)
    assert bplan.exec_ptr(d.data_ptr(), d.numel(), peer.row_ptr(f0), st) == f1 - f0
    peer.finish()
    if rank == 0:
        got = torch.empty((frames, 516), device="cuda")
        peer.copy_to(got.data_ptr())
        assert torch.equal(got[:, :513], ref)
        print("PEER_BULK_OK")
    dist.barrier()
    if rank != 0:
        peer.close()
    dist.barrier()
    if rank == 0:
        peer.close()
dist.barrier()
dist.destroy_process_group()
'''


def test_sharded_stft_gather_nccl(built_lib, tmp_path):
    import torch
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29700 + os.getpid() % 200), str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "SHARDED_OK" in r.stdout and "PEER_DIRECT_OK" in r.stdout and "PEER_BULK_OK" in r.stdout


_DIST_FFT_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from kopek_b200.sharded import DistributedFFT, dist_fft_layout
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
if world not in (2, 4, 8):
    world_ok = False
for lg in (18, 22):
    n = world << lg                                            # local transforms of 2^18 and 2^22 points
    g = torch.Generator().manual_seed(lg)                      # the same signal on every rank
    x = torch.view_as_complex(torch.randn(n, 2, generator=g))
    x[5 * world + 1] += 300.0                                  # something that is not noise
    src, dst = dist_fft_layout(n, world, rank)
    mine = x[torch.from_numpy(src)].cuda().contiguous()
    out = torch.full((world, n // world // world), float("nan"), dtype=torch.complex64, device="cuda")
    f = DistributedFFT(n, local)
    f.forward(mine.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    ref = torch.fft.fft(x.to(torch.complex128))                 # f64 on the host: the whole transform
    want = ref[torch.from_numpy(dst.reshape(-1))].reshape(dst.shape)
    err = ((out.cpu().to(torch.complex128) - want).abs().max() / ref.abs().max()).item()
    assert err <= 1e-5 * np.log2(n), (lg, err)
    # second call on new data reuses the mapped buffers
    y = torch.view_as_complex(torch.randn(n, 2, generator=g))
    mine2 = y[torch.from_numpy(src)].cuda().contiguous()
    f.forward(mine2.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    want2 = torch.fft.fft(y.to(torch.complex128))[torch.from_numpy(dst.reshape(-1))].reshape(dst.shape)
    err2 = ((out.cpu().to(torch.complex128) - want2).abs().max() / want2.abs().max()).item()
    assert err2 <= 1e-5 * np.log2(n), (lg, err2)
    f.close()
    if rank == 0:
        print("DIST_FFT_OK", world, n, f"{err:.2e}")
dist.barrier()
dist.destroy_process_group()
'''


def test_one_large_transform_over_all_gpus(built_lib, tmp_path):
    """a single transform of world * 2^22 points, the inter-rank butterfly done by ONE kernel per
    rank that reads the peers' partial spectra over NVLink; checked against an f64 transform of the whole signal."""
    import torch
    world = torch.cuda.device_count()
    world = 8 if world >= 8 else 4 if world >= 4 else 2 if world >= 2 else 0
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    script = tmp_path / "dist_fft_worker.py"
    script.write_text(_DIST_FFT_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29900 + os.getpid() % 90), str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("DIST_FFT_OK") == 2
