"""CPU: host-side logic -- register butterflies on the host, frame sharding, gloo gather."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_register_butterflies_on_host(tmp_path):
    exe = str(tmp_path / "test_butterflies")
    src = os.path.join(ROOT, "tests", "cpp", "test_butterflies.cu")
    subprocess.run(["nvcc", "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets",
                    "-o", exe, src], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_frame_and_sample_ranges():
    from kopek_b200.sharded import frame_range, sample_range
    for frames in (0, 1, 7, 5622, 1000000):
        for world in (1, 2, 3, 4, 8):
            spans = [frame_range(r, world, frames) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == frames
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    # STFT halo: neighbours overlap by n - hop samples
    n, hop, ns = 2048, 512, 2880000
    s0 = sample_range(0, 2, ns, n, hop)
    s1 = sample_range(1, 2, ns, n, hop)
    assert s0[2] == 0 and s0[3] == s1[2] and s1[3] == 5622
    assert s0[1] - s1[0] == n - hop
    assert s1[1] == (5622 - 1) * hop + n <= ns
    assert sample_range(0, 4, 100, 256, 64) == (0, 0, 0, 0)


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
import oracle
from kopek_b200.sharded import sample_range, gather_spectra
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
n, hop = 256, 64
x = np.random.default_rng(3).standard_normal(5000).astype(np.float32)   # same buffer on every rank
s0, s1, f0, f1 = sample_range(rank, world, x.size, n, hop)
# the shard's spectra (oracle stands in for the GPU kernel in this CPU test of the plumbing)
local = torch.from_numpy(oracle.stft_mag(x[s0:s1], n, hop, oracle.WIN_HANN)) if f1 > f0 else torch.empty((0, n // 2 + 1))
assert local.shape[0] == f1 - f0
frames = oracle.stft_frames(x.size, n, hop)
full = gather_spectra(local, frames, dst=0)
if rank == 0:
    ref = oracle.stft_mag(x, n, hop, oracle.WIN_HANN)
    assert full.shape == ref.shape
    assert np.array_equal(full.numpy(), ref)      # sharded result is bit-identical to the unsharded one
    print("GATHER_OK")
else:
    assert full is None
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_gather_gloo(tmp_path, oracle_mod, world):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = 29600 + world + (os.getpid() % 200)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "GATHER_OK" in outs[0]


def test_dist_fft_layout_partitions_the_transform():
    """Host logic of DistributedFFT: the cyclic input slices partition the signal, the output blocks the spectrum, and
    the decimation identity the combine kernel implements holds (numpy, f64)."""
    from kopek_b200.sharded import dist_fft_layout
    rng = np.random.default_rng(3)
    for world, n in ((2, 64), (4, 256), (8, 1024)):
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        X = np.fft.fft(x)
        L = n // world
        seen_in, seen_out = np.zeros(n, int), np.zeros(n, int)
        F = [np.fft.fft(x[dist_fft_layout(n, world, r)[0]]) for r in range(world)]       # local transforms
        for p in range(world):
            src, dst = dist_fft_layout(n, world, p)
            seen_in[src] += 1
            seen_out[dst.reshape(-1)] += 1
            kp = p * (L // world) + np.arange(L // world)
            v = np.stack([F[r][kp] * np.exp(-2j * np.pi * r * kp / n) for r in range(world)])  # W_n^{r k'} F_r[k']
            out = np.fft.fft(v, axis=0)                                                        # radix-world butterfly
            assert np.abs(out - X[dst]).max() < 1e-9 * np.abs(X).max()
        assert (seen_in == 1).all() and (seen_out == 1).all()
