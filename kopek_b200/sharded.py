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
        ops = []
        for r in range(world):
            f0, f1 = frame_range(r, world, n_frames_total)
            if r == dst:
                full[f0:f1].copy_(local)
            elif f1 > f0:
                ops.append(dist.P2POp(dist.irecv, full[f0:f1], r, group))
        if ops:                                   # one grouped launch (ncclGroupStart/End): the receives overlap
            for q in dist.batch_isend_irecv(ops):
                q.wait()
        return full
    if local.shape[0] > 0:
        for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, local.contiguous(), dst, group)]):
            q.wait()
    return None


class PeerSpectrogram:
    """One [n_frames_total, pitch] spectrogram on `dst`'s GPU that every rank's kernel writes into directly.

    `dst` allocates the buffer (kopek_device_alloc) and broadcasts its CUDA IPC handle; the other ranks map it
    (kopek_ipc_open) and pass `row_ptr(f0)` as d_out of StftPlan.exec_ptr, so K1/K1W store their bins straight
    into the owner's HBM over NVLink -- transform and transfer are one kernel, no gather afterwards.
    Call `finish()` (device sync + barrier) before the owner reads; `copy_to(ptr)` on the owner.  `close()` is not
    collective: the importing ranks close first, then (after a barrier of the caller's) the owner -- CUDA IPC requires
    every mapping to be gone before the exporting process frees the allocation.
    """

    def __init__(self, n_frames_total: int, pitch: int, elem_bytes: int, local_device: int, dst: int = 0, group=None):
        import ctypes as C

        import torch.distributed as dist

        from . import _lib
        self._lib, self._check, self._C = _lib.load(), _lib.check, C
        self.frames, self.pitch, self.elem_bytes, self.device = n_frames_total, pitch, elem_bytes, local_device
        self.group, self.dst = group, dst
        self.rank = dist.get_rank(group)
        self.nbytes = n_frames_total * pitch * elem_bytes
        self._ptr = C.c_void_p()
        handle = [None]
        if self.rank == dst:
            self._check(self._lib.kopek_device_alloc(C.byref(self._ptr), self.nbytes, local_device))
            raw = (C.c_ubyte * 64)()
            self._check(self._lib.kopek_ipc_export(self._ptr, raw))
            handle = [bytes(raw)]
        dist.broadcast_object_list(handle, src=dst, group=group)
        if self.rank != dst:
            raw = (C.c_ubyte * 64).from_buffer_copy(handle[0])
            self._check(self._lib.kopek_ipc_open(raw, local_device, C.byref(self._ptr)))

    def row_ptr(self, frame: int) -> int:
        return self._ptr.value + frame * self.pitch * self.elem_bytes

    def finish(self) -> None:
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize(self.device)        # this rank's peer stores have left the SMs and are visible
        dist.barrier(group=self.group)

    def copy_to(self, dst_ptr: int) -> None:
        """Owner only: copy the finished spectrogram to a host or device pointer."""
        assert self.rank == self.dst
        self._check(self._lib.kopek_copy(dst_ptr, self._ptr, self.nbytes, self.device))

    def close(self) -> None:
        if self._ptr:
            if self.rank == self.dst:
                self._lib.kopek_device_free(self._ptr, self.device)
            else:
                self._lib.kopek_ipc_close(self._ptr, self.device)
            self._ptr = self._C.c_void_p()


# ---- one large transform over the ranks -----------------------------------------------------
def dist_fft_layout(n: int, world: int, rank: int):
    """Index bookkeeping of DistributedFFT: (input indices of this rank, output indices of this rank as [world, block]).

    Input: the cyclic slice x[j * world + rank], j < n / world.  Output: out[q][i] = X[(rank * block + i) + L * q] with
    L = n / world, block = L / world."""
    import numpy as np
    L = n // world
    block = L // world
    src = np.arange(L, dtype=np.int64) * world + rank
    dst = (rank * block + np.arange(block, dtype=np.int64))[None, :] + L * np.arange(world, dtype=np.int64)[:, None]
    return src, dst


class DistributedFFT:
    """One complex-f32 transform of n = world * L points over `world` GPUs (one process per GPU, world in {2, 4, 8}).

    Rank r holds the cyclic slice x[j * world + r].  `forward` runs the four-step transform on the local slice
    (kopek_fft_large_c2c_f32: no traffic), then ONE kernel per rank finishes its block of bins, reading the other ranks'
    partial spectra straight out of their HBM over NVLink (CUDA-IPC peer mappings; kopek_fft_dist_combine) while it
    applies the inter-rank twiddle and the radix-`world` butterfly -- the exchange and the arithmetic are one kernel.
    Output layout: see dist_fft_layout.  L must be a power of two in [2^16, 2^24]."""

    def __init__(self, n: int, local_device: int, group=None):
        import ctypes as C

        import torch.distributed as dist

        from . import _lib
        self._lib, self._check, self._C = _lib.load(), _lib.check, C
        self.group, self.device = group, local_device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world not in (2, 4, 8):
            raise ValueError("DistributedFFT needs 2, 4 or 8 ranks")
        if n % (self.world * self.world):
            raise ValueError("n must be a multiple of world^2")
        self.n, self.L = n, n // self.world
        self.block = self.L // self.world
        # this rank's partial spectrum F_r lives in an IPC-exportable allocation; everybody maps everybody's
        self._own = C.c_void_p()
        self._check(self._lib.kopek_device_alloc(C.byref(self._own), self.L * 8, local_device))
        raw = (C.c_ubyte * 64)()
        self._check(self._lib.kopek_ipc_export(self._own, raw))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(raw), group=group)
        self._peers = []
        for r in range(self.world):
            if r == self.rank:
                self._peers.append(C.c_void_p(self._own.value))
            else:
                ptr = C.c_void_p()
                buf = (C.c_ubyte * 64).from_buffer_copy(handles[r])
                self._check(self._lib.kopek_ipc_open(buf, local_device, C.byref(ptr)))
                self._peers.append(ptr)
        self._table = (C.c_void_p * self.world)(*[p.value for p in self._peers])

    def forward(self, d_in: int, d_out: int, stream: int = 0) -> None:
        """d_in: L complex64 (the cyclic slice), d_out: [world, block] complex64; both device pointers of this rank."""
        import torch
        import torch.distributed as dist
        self._check(self._lib.kopek_fft_large_c2c_f32(d_in, self._own, self.L, self.device, stream or None))
        torch.cuda.synchronize(self.device)          # F_r is complete in this rank's HBM ...
        dist.barrier(group=self.group)               # ... and in everybody else's
        self._check(self._lib.kopek_fft_dist_combine(self._table, self.world, self.rank, self.L, d_out, self.device,
                                                     stream or None))
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)               # nobody overwrites its F_r while a peer still reads it

    def combine_only(self, d_out: int, stream: int = 0) -> None:
        """The exchange kernel alone (timing): F_r must be in place on every rank."""
        self._check(self._lib.kopek_fft_dist_combine(self._table, self.world, self.rank, self.L, d_out, self.device,
                                                     stream or None))

    def close(self) -> None:
        import torch.distributed as dist
        dist.barrier(group=self.group)
        for r, p in enumerate(self._peers):
            if r != self.rank and p:
                self._lib.kopek_ipc_close(p, self.device)
        self._peers = []
        dist.barrier(group=self.group)               # every importer has unmapped before an owner frees (CUDA IPC rule)
        if self._own:
            self._lib.kopek_device_free(self._own, self.device)
            self._own = self._C.c_void_p()


# ---- one process, R devices: the C ABI's own multi-GPU entry points (include/kopek_cuda.h section 4) ---------------
class MultiGpuStft:
    """kopek_mgpu_*: frames of one buffer sharded over `devices` from ONE process, no torch on the data path.

    `exec_host(samples)` is the call a reference-side caller makes (host buffer in, [frames, bins] spectrogram out);
    `exec_ptrs` / `exec_gathered_ptr` / `gather_ptrs` work on device pointers; `sync()` returns the device time (ms,
    max over the devices) of everything enqueued since the previous sync."""

    def __init__(self, devices, n: int, hop: int | None = None, window: int = 0, output: int = 1, bins: int = 0,
                 out_pitch: int = 0):
        import ctypes as C

        from . import _lib
        self._C, self._lib, self._check = C, _lib.load(), _lib.check
        self.devices = [int(d) for d in devices]
        self.n, self.hop = n, (n if hop is None else hop)
        self._h = C.c_void_p()
        arr = (C.c_int32 * len(self.devices))(*self.devices)
        self._check(self._lib.kopek_mgpu_create(C.byref(self._h), arr, len(self.devices), n, self.hop, window, output,
                                                bins, out_pitch))
        self.bins = int(self._lib.kopek_mgpu_bins(self._h))
        self.pitch = int(self._lib.kopek_mgpu_out_pitch(self._h))
        self.elem_bytes = int(self._lib.kopek_mgpu_out_elem_bytes(self._h))

    def frames(self, n_samples: int) -> int:
        return 0 if n_samples < self.n else 1 + (n_samples - self.n) // self.hop

    def shard(self, n_samples: int, rank: int):
        """(sample_begin, sample_count, frame_begin, frame_count) of device `rank`."""
        C = self._C
        v = [C.c_uint64(0) for _ in range(4)]
        self._check(self._lib.kopek_mgpu_shard(self._h, n_samples, rank, *[C.byref(x) for x in v]))
        return tuple(int(x.value) for x in v)

    def _ptr_array(self, ptrs):
        return (self._C.c_void_p * len(ptrs))(*[int(p) for p in ptrs])

    def _u64_array(self, vals):
        return (self._C.c_uint64 * len(vals))(*[int(v) for v in vals])

    def exec_ptrs(self, d_samples, n_samples, d_out):
        got = self._u64_array([0] * len(self.devices))
        self._check(self._lib.kopek_mgpu_exec(self._h, self._ptr_array(d_samples), self._u64_array(n_samples),
                                              self._ptr_array(d_out), got))
        return [int(g) for g in got]

    def exec_gathered_ptr(self, d_samples, n_samples, d_dst: int) -> int:
        total = self._C.c_uint64(0)
        self._check(self._lib.kopek_mgpu_exec_gathered(self._h, self._ptr_array(d_samples), self._u64_array(n_samples),
                                                       d_dst, self._C.byref(total)))
        return int(total.value)

    def gather_ptrs(self, d_out, n_frames, d_dst: int, mode: int = 0) -> None:
        self._check(self._lib.kopek_mgpu_gather(self._h, self._ptr_array(d_out), self._u64_array(n_frames), d_dst, mode))

    def sync(self) -> float:
        ms = self._C.c_float(0.0)
        self._check(self._lib.kopek_mgpu_sync(self._h, self._C.byref(ms)))
        return float(ms.value)

    def exec_host_ptr(self, h_samples: int, n_samples: int, h_out: int) -> int:
        nf = self._C.c_uint64(0)
        self._check(self._lib.kopek_mgpu_exec_host(self._h, h_samples, n_samples, h_out, self._C.byref(nf)))
        return int(nf.value)

    def exec_host(self, samples):
        import numpy as np
        s = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1)
        frames = self.frames(s.size)
        out = np.empty((frames, self.pitch), dtype=np.complex64 if self.elem_bytes == 8 else np.float32)
        if frames:
            self.exec_host_ptr(s.ctypes.data, s.size, out.ctypes.data)
        return out[:, :self.bins]

    def close(self) -> None:
        if self._h:
            self._lib.kopek_mgpu_destroy(self._h)
            self._h = self._C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
