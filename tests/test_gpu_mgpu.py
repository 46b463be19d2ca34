"""GPU: the C ABI's own multi-GPU entry points (kopek_mgpu_*, include/kopek_cuda.h section 4) from ONE process.

Every check is "bit-identical to the one-device result" (frames are independent); the one-device result
itself is held to the oracle in test_gpu_stft.py.  With a single visible GPU the same code runs on the device list [0]
(and the sharding arithmetic is still exercised against kopek_b200.sharded); with 2+ GPUs the gathers cross NVLink.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _devices(torch, at_most=8):
    return list(range(min(torch.cuda.device_count(), at_most)))


def test_shard_ranges_match_the_python_mirror(built_lib):
    import torch

    from kopek_b200.sharded import MultiGpuStft, sample_range
    devs = _devices(torch)
    for n, hop, ns in ((1024, 256, 1_000_003), (2048, 512, 2_880_000), (256, 256, 1000 * 256), (4096, 1024, 4095)):
        with MultiGpuStft(devs, n, hop) as m:
            total = 0
            for r in range(len(devs)):
                s0, cnt, f0, nf = m.shard(ns, r)
                ps0, ps1, pf0, pf1 = sample_range(r, len(devs), ns, n, hop)
                assert (f0, nf) == (pf0, pf1 - pf0)
                assert (s0, cnt) == ((ps0, ps1 - ps0) if nf else (0, 0))
                total += nf
            assert total == m.frames(ns)


@pytest.mark.parametrize("n,hop,window,output,pitch,bins", [
    (1024, 256, 1, 1, 0, 0), (1024, 1024, 1, 1, 516, 0), (2048, 512, 1, 3, 0, 0), (256, 256, 0, 1, 0, 0),
    (4096, 1024, 2, 2, 0, 0), (1024, 512, 0, 0, 0, 0),
    (1024, 1001, 1, 1, 0, 1), (4096, 1023, 1, 1, 0, 1), (512, 77, 2, 0, 0, 1), (1024, 999, 1, 4, 0, 0)])
def test_exec_host_is_bit_identical_to_one_device(built_lib, n, hop, window, output, pitch, bins):
    """Shards start wherever floor(r F / R) * hop falls: with odd hops a device's sub-buffer is only 4-byte aligned, so
    the same frame may take the 32-bit-load instantiation on one device count and the vector one on another -- the
    two are bit-identical (tests/test_gpu_stft.py), and so is the sharded result."""
    import torch

    from kopek_b200 import StftPlan, signals
    from kopek_b200.sharded import MultiGpuStft
    devs = _devices(torch)
    x = signals.chirp(700_001, 48000.0, 20.0, 20000.0) + 0.01 * signals.white_noise(700_001, seed=n)
    with StftPlan(n, hop, window, output, bins=bins, out_pitch=pitch, device=0) as plan:
        ref = plan.exec_host(x)
    with MultiGpuStft(devs, n, hop, window, output, bins=bins, out_pitch=pitch) as m:
        got = m.exec_host(x)
        assert got.shape == ref.shape and got.shape[0] == m.frames(x.size)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
        again = m.exec_host(x[: 50 * hop + n])                  # shorter buffer through the same context
        assert np.array_equal(again.view(np.uint32), ref[:51].view(np.uint32))
        assert m.exec_host(x[: n - 1]).shape[0] == 0             # fewer samples than one frame


@pytest.mark.parametrize("pitch", [0, 516])
@pytest.mark.parametrize("mode", ["gathered", "copy", "nccl"])
def test_device_shards_and_gather_to_device0(built_lib, pitch, mode):
    """Device-resident shards -> ONE spectrogram on devices[0]: fused peer-direct stores, copy-engine gather, NCCL."""
    import torch

    from kopek_b200 import StftPlan, _lib, signals
    from kopek_b200.sharded import MultiGpuStft
    devs = _devices(torch)
    n, hop = 1024, 256
    x = signals.chirp(1_000_003, 48000.0, 20.0, 20000.0)
    with StftPlan(n, hop, 1, 1, out_pitch=pitch, device=0) as plan:
        frames = plan.frames(x.size)
        dx = torch.from_numpy(x).cuda(0)
        ref = torch.empty((frames, plan.pitch), device="cuda:0")
        plan.exec_ptr(dx.data_ptr(), dx.numel(), ref.data_ptr(), torch.cuda.current_stream(0).cuda_stream)
        torch.cuda.synchronize(0)
        bins = plan.bins
    with MultiGpuStft(devs, n, hop, 1, 1, out_pitch=pitch) as m:
        ins, outs, counts, nfs = [], [], [], []
        for r, dev in enumerate(devs):
            s0, cnt, f0, nf = m.shard(x.size, r)
            ins.append(torch.from_numpy(x[s0:s0 + cnt]).cuda(dev))
            outs.append(torch.empty((max(nf, 1), m.pitch), device=f"cuda:{dev}"))
            counts.append(cnt)
            nfs.append(nf)
        for d in devs:
            torch.cuda.synchronize(d)
        full = torch.full((frames, m.pitch), float("nan"), device="cuda:0")
        torch.cuda.synchronize(0)
        if mode == "gathered":
            total = m.exec_gathered_ptr([t.data_ptr() for t in ins], counts, full.data_ptr())
            assert total == frames
        else:
            got = m.exec_ptrs([t.data_ptr() for t in ins], counts, [t.data_ptr() for t in outs])
            assert got == nfs
            try:
                m.gather_ptrs([t.data_ptr() for t in outs], nfs, full.data_ptr(),
                              _lib.GATHER_NCCL if mode == "nccl" else _lib.GATHER_COPY)
            except _lib.KopekError as e:
                if mode == "nccl" and e.status == _lib.ERR_UNSUPPORTED:
                    pytest.skip("libnccl.so.2 not loadable on this box")
                raise
        ms = m.sync()
        assert ms > 0.0
        assert torch.equal(full[:, :bins], ref[:, :bins]), mode


def test_mgpu_argument_errors(built_lib):
    from kopek_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    two = (C.c_int32 * 2)(0, 0)
    assert lib.kopek_mgpu_create(C.byref(h), two, 2, 1024, 1024, 0, 1, 0, 0) == _lib.ERR_INVALID      # device twice
    assert b"twice" in lib.kopek_last_error()
    assert lib.kopek_mgpu_create(C.byref(h), two, 0, 1024, 1024, 0, 1, 0, 0) == _lib.ERR_INVALID      # empty list
    one = (C.c_int32 * 1)(0)
    assert lib.kopek_mgpu_create(C.byref(h), one, 1, 1000, 1000, 0, 1, 0, 0) == _lib.ERR_UNSUPPORTED  # frame size
    bad = (C.c_int32 * 1)(99)
    assert lib.kopek_mgpu_create(C.byref(h), bad, 1, 1024, 1024, 0, 1, 0, 0) == _lib.ERR_INVALID      # no such device
    assert lib.kopek_mgpu_destroy(None) == _lib.OK


def test_failed_exec_leaves_the_context_usable(built_lib):
    """An entry point that fails after the timing window was opened (here: a NULL shard pointer) must leave the
    context in a state kopek_mgpu_sync accepts, and the next valid call must give the usual result."""
    import torch

    from kopek_b200 import StftPlan, _lib, signals
    from kopek_b200.sharded import MultiGpuStft
    devs = _devices(torch)
    n, hop = 1024, 512
    x = signals.white_noise(200_000, seed=11)
    with StftPlan(n, hop, 1, 1, device=0) as plan:
        ref = plan.exec_host(x)
    with MultiGpuStft(devs, n, hop, 1, 1) as m:
        lib = _lib.load()
        R = len(devs)
        ins = (C.c_void_p * R)(*([None] * R))                      # NULL sample pointers with a non-zero count
        cnt = (C.c_uint64 * R)(*([4096] * R))
        outs = (C.c_void_p * R)(*([None] * R))
        assert lib.kopek_mgpu_exec(m._h, ins, cnt, outs, None) == _lib.ERR_INVALID
        ms = C.c_float(-1.0)
        assert lib.kopek_mgpu_sync(m._h, C.byref(ms)) == _lib.OK    # drains, reports, resets the window
        got = m.exec_host(x)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
        assert lib.kopek_mgpu_sync(m._h, C.byref(ms)) == _lib.OK


def test_custom_window_on_every_device(built_lib):
    """kopek_mgpu_set_custom_window uploads the coefficients to every device's plan."""
    import torch

    from kopek_b200 import StftPlan, WINDOW_CUSTOM, _lib, signals
    from kopek_b200.sharded import MultiGpuStft
    devs = _devices(torch)
    n, hop = 2048, 700
    w = np.kaiser(n, 6.0).astype(np.float32)
    x = signals.white_noise(300_000, seed=5)
    with StftPlan(n, hop, WINDOW_CUSTOM, 1, custom_window=w, device=0) as plan:
        ref = plan.exec_host(x)
    with MultiGpuStft(devs, n, hop, WINDOW_CUSTOM, 1) as m:
        _lib.check(_lib.load().kopek_mgpu_set_custom_window(m._h, w.ctypes.data))
        got = m.exec_host(x)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
