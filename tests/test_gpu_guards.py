"""GPU: out-of-bounds canaries (compute-sanitizer is closed on this pool): every kernel writes into the middle of a
sentinel-filled allocation and must leave the guard zones before and after its output untouched."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GUARD = 4096          # elements on each side
SENT = -12345.0


def _guarded(n_elems, dtype):
    import torch
    buf = torch.full((n_elems + 2 * GUARD,), SENT, device="cuda", dtype=dtype)
    return buf, buf[GUARD:GUARD + n_elems]


def _check(buf, n_elems):
    import torch
    torch.cuda.synchronize()
    lo, hi = buf[:GUARD], buf[GUARD + n_elems:]
    ref = torch.full_like(lo, SENT)
    assert torch.equal(lo, ref) and torch.equal(hi, ref), "kernel wrote outside its output"
    assert not bool((buf[GUARD:GUARD + n_elems] == SENT).any()), "kernel left output elements unwritten"


@pytest.mark.parametrize("n,hop,window,output,bins", [
    (1024, 1024, 1, 1, 0), (1024, 1024, 0, 3, 0), (1024, 256, 1, 0, 0), (1024, 1024, 0, 4, 0), (1024, 1024, 1, 5, 0),
    (1024, 333, 1, 1, 0), (256, 256, 0, 1, 0), (512, 128, 1, 2, 0), (2048, 512, 1, 3, 0), (4096, 1024, 1, 1, 0),
    (4096, 4096, 0, 0, 0),
    # round 2: FULL rows (mirror stores) and 32-bit-load instantiations of every lane-group kernel
    (1024, 1024, 1, 1, 1), (1024, 1023, 0, 0, 1), (256, 255, 1, 1, 1), (512, 512, 0, 3, 1), (512, 511, 1, 1, 0),
    (2048, 2048, 1, 0, 1), (2048, 2047, 2, 1, 1), (4096, 4095, 1, 1, 1), (4096, 1026, 0, 2, 0), (1024, 1021, 1, 4, 0),
    (2048, 512, 3, 1, 1), (4096, 4097, 3, 3, 0)])
def test_stft_kernels_stay_in_bounds(built_lib, n, hop, window, output, bins):
    import torch
    from kopek_b200 import StftPlan
    frames = 1237                                     # not a multiple of frames-per-CTA or of the resident warps
    ns = (frames - 1) * hop + n
    xin, x = _guarded(ns, torch.float32)
    x.copy_(torch.randn(ns, device="cuda") + 2.0)     # strictly non-zero spectra so "unwritten" is detectable
    custom = np.hanning(n).astype(np.float32) + 0.1 if window == 3 else None
    with StftPlan(n, hop, window, output, bins=bins, custom_window=custom) as plan:
        ne = frames * plan.pitch
        obuf, out = _guarded(ne, torch.complex64 if output == 0 else torch.float32)
        assert plan.exec_ptr(x.data_ptr(), ns, out.data_ptr(), torch.cuda.current_stream().cuda_stream) == frames
        _check(obuf, ne)
    # the input guards are untouched as well (kernels only read, TMA prefetch must not run past the last frame)
    torch.cuda.synchronize()
    assert bool((xin[:GUARD] == SENT).all()) and bool((xin[GUARD + ns:] == SENT).all())


@pytest.mark.parametrize("log_n", [16, 19, 22, 23, 24])      # 23, 24: the radix-64 pass 2
def test_large_fft_stays_in_bounds(built_lib, log_n):
    import torch
    from kopek_b200 import fft
    n = 1 << log_n
    a = torch.randn(n, dtype=torch.complex64, device="cuda") + 1.0
    obuf, out = _guarded(n, torch.complex64)
    fft.fft_large_device(a.data_ptr(), out.data_ptr(), n, 0, torch.cuda.current_stream().cuda_stream)
    _check(obuf, n)


def test_drop_in_device_stays_in_bounds(built_lib):
    import torch
    from kopek_b200 import _lib
    batch, n_in = 37, 1000
    x = torch.randn((batch, n_in), dtype=torch.complex128, device="cuda") + 1.0
    obuf, out = _guarded(batch * 1024, torch.complex128)
    _lib.check(_lib.load().kopek_fft_c2c_f64_device(x.data_ptr(), n_in, batch, out.data_ptr(), 0,
                                                    torch.cuda.current_stream().cuda_stream))
    _check(obuf, batch * 1024)


def test_exec_host_drains_every_chunk_before_reporting_a_failure(built_lib, monkeypatch):
    """Fault injection (KOPEK_TEST_FAIL_CHUNK): the enqueue of chunk 2 fails while chunks 0 and 1 are in flight on
    the other slot streams.  The call must report the failure AND have drained those chunks: their rows are complete
    in the caller's buffer when it returns (the caller may free it immediately), later rows are untouched, and the
    plan stays usable."""
    from kopek_b200 import KopekError, StftPlan, signals
    n, frames = 1024, 4096
    x = signals.sine_frames(frames, n)
    monkeypatch.setenv("KOPEK_HOST_CHUNK_MB", "1")               # 256 frames per chunk -> 16 chunks
    with StftPlan(n, n, 1, 1) as plan:
        ref = plan.exec_host(x).copy()
        out = np.full((frames, plan.pitch), SENT, dtype=np.float32)
        monkeypatch.setenv("KOPEK_TEST_FAIL_CHUNK", "2")
        with pytest.raises(KopekError, match="injected"):
            plan.exec_host_ptr(x.ctypes.data, x.size, out.ctypes.data)
        assert plan.last_launches == 2
        assert np.array_equal(out[:512], ref[:512]), "chunks enqueued before the failure were not drained"
        assert (out[512:] == SENT).all()
        monkeypatch.delenv("KOPEK_TEST_FAIL_CHUNK")
        assert np.array_equal(plan.exec_host(x), ref)


def test_exec_host_chunks_are_sized_by_the_busier_direction(built_lib):
    """hop = 1 writes 2049 x 4 bytes per 4 bytes read; slots must be sized by the output, not the input
    (the first version asked for ~34 GB per slot here)."""
    import torch
    from kopek_b200 import StftPlan, signals
    n, hop = 4096, 1
    x = signals.white_noise(n + 20000, seed=3)
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    with StftPlan(n, hop, 1, 1) as plan:
        got = plan.exec_host(x)
        used = free0 - torch.cuda.mem_get_info()[0]
        assert used < (1 << 30), f"chunk slots took {used >> 20} MiB"
        with StftPlan(n, 1024, 1, 1) as coarse:
            ref = coarse.exec_host(x)
        # every 1024th frame is a frame of the coarse plan (another kernel: same values to rounding)
        assert got.shape[0] == 20001 and ref.shape[0] == 20
        assert np.abs(got[::1024][:20] - ref).max() <= 1e-5 * 12 * ref.max()


@pytest.mark.parametrize("n,hop,output", [(1024, 1024, 1), (4096, 4096, 1), (256, 256, 0)])
def test_sample_and_byte_offsets_beyond_32_bits(built_lib, n, hop, output):
    """Maximum sizes: one call over MORE than 2^32 samples (17 GB of f32) whose output crosses 2^32 bytes (and, for the
    complex rows, 2^32 elements too).  The signal is zero except for a handful of frames around the 2^32-sample and
    2^32-byte marks and at the very end; those rows must equal the rows of the same frames transformed on their
    own, every row between them must be exactly zero, and nothing may land outside the output."""
    import torch

    from kopek_b200 import StftPlan, signals
    free, _ = torch.cuda.mem_get_info()
    frames = (1 << 32) // hop + 37
    ns = (frames - 1) * hop + n
    with StftPlan(n, hop, 1, output, device=0) as plan:
        row = plan.pitch * (2 if output == 0 else 1)                  # f32 values per row
        need = ns * 4 + frames * row * 4 + (2 << 30)
        if free < need:
            pytest.skip(f"needs {need >> 30} GiB of device memory")
        x = torch.zeros(ns, device="cuda", dtype=torch.float32)
        # frames of interest: around sample 2^32, around output byte 2^32, and the last ones
        f_byte = (1 << 32) // (row * 4)
        marks = sorted({0, 1, f_byte - 1, f_byte, f_byte + 1, (1 << 32) // hop - 1, (1 << 32) // hop, frames - 2, frames - 1})
        tone = torch.from_numpy(signals.sine_frames(len(marks), n)).cuda().view(len(marks), n)
        for i, f in enumerate(marks):
            x[f * hop:f * hop + n] = tone[i]
        out = torch.full((frames * row + 2 * GUARD,), SENT, device="cuda", dtype=torch.float32)
        got = plan.exec_ptr(x.data_ptr(), ns, out[GUARD:].data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert got == frames
        small = torch.empty((len(marks), row), device="cuda", dtype=torch.float32)
        plan.exec_ptr(tone.data_ptr(), tone.numel(), small.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        used = plan.bins * (2 if output == 0 else 1)
    body = out[GUARD:GUARD + frames * row].view(frames, row)
    assert torch.equal(out[:GUARD], torch.full_like(out[:GUARD], SENT))
    assert torch.equal(out[GUARD + frames * row:], torch.full_like(out[:GUARD], SENT))
    for i, f in enumerate(marks):
        assert torch.equal(body[f, :used].view(torch.int32), small[i, :used].view(torch.int32)), f"frame {f}"
    # everything else is the transform of silence: exactly zero (checked in slabs to bound the temporaries)
    keep = torch.ones(frames, dtype=torch.bool, device="cuda")
    keep[torch.tensor(marks, device="cuda")] = False
    step = 1 << 18
    for f0 in range(0, frames, step):
        blk = body[f0:f0 + step, :used]
        assert not bool((blk[keep[f0:f0 + step]] != 0).any()), f"non-zero row in frames {f0}.."
