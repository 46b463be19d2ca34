"""Frame-range sharding of a batched STFT over one process per GPU.

Frames are independent (the reference's fft is a pure function, src/fft.rs:6),
so the transform needs no data-path collective: rank r of R owns the contiguous
frame range [floor(r F / R), floor((r+1) F / R)) and reads the sample range that
covers it (neighbouring ranks duplicate the (n - hop)-sample halo when frames
overlap).  Spectra stay sharded unless the caller asks for one spectrogram, in
which case `gather_spectra` collects the blocks on rank 0 with torch.distributed
(NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Tuple


def frame_range(rank: int, world: int, n_frames: int) -> Tuple[int, int]:
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return (rank * n_frames) // world, ((rank + 1) * n_frames) // world


def sample_range(rank: int, world: int, n_samples: int, n: int, hop: int) -> Tuple[int, int, int, int]:
    """(sample_begin, sample_end, frame_begin, frame_end) of this rank's shard."""
    n_frames = 0 if n_samples < n else 1 + (n_samples - n) // hop
    f0, f1 = frame_range(rank, world, n_frames)
    if f1 <= f0:
        return 0, 0, f0, f1
    return f0 * hop, (f1 - 1) * hop + n, f0, f1


def gather_spectra(local, n_frames_total: int, dst: int = 0, group=None):
    """Gather per-rank [frames_r, bins] tensors into one [F, bins] tensor on `dst`.

    Ranks hold different row counts, so this is a grouped send/recv (NCCL has no
    native gatherv): `dst` posts one irecv per peer straight into the row slice
    of the result, the peers isend their block.  Returns the full tensor on
    `dst`, None elsewhere.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return local
    if rank == dst:
        full = torch.empty((n_frames_total, local.shape[1]), dtype=local.dtype, device=local.device)
        reqs = []
        for r in range(world):
            f0, f1 = frame_range(r, world, n_frames_total)
            if r == dst:
                full[f0:f1].copy_(local)
            elif f1 > f0:
                reqs.append(dist.irecv(full[f0:f1], src=r, group=group))
        for q in reqs:
            q.wait()
        return full
    if local.shape[0] > 0:
        dist.isend(local.contiguous(), dst=dst, group=group).wait()
    return None
