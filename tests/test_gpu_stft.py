"""GPU: the fused window -> FFT -> |X|/dB kernel (K1) against the oracle, through the C ABI.

Tolerance (north_star): per frame max_k |X_gpu - X_ref| <= 1e-5*log2(N) * max_k |X_ref| against the
f32 AND the f64 instantiation of the restated recursion; peak bins equal.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from util import frame_rel_err, peaks_match, tol

pytestmark = pytest.mark.gpu

SIZES = [256, 512, 1024, 2048, 4096]


def _run(plan, x, pitch_view=True):
    import torch
    d_in = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda()
    frames = plan.frames(d_in.numel())
    out = torch.full((max(frames, 1), plan.pitch), float("nan"), device="cuda",
                     dtype=torch.complex64 if plan.output == 0 else torch.float32)
    got = plan.exec_ptr(d_in.data_ptr(), d_in.numel(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert got == frames
    return out[:frames].cpu().numpy()


def _signals(n, frames):
    from kopek_b200 import signals
    total = n * frames
    sine = signals.sine_frames(frames, n, noise=0.0)
    return {
        "sine": sine,
        "sine+noise": signals.sine_frames(frames, n, noise=0.1, seed=1234),
        "chirp": signals.chirp(total, 48000.0, 20.0, 20000.0),
        "white": signals.white_noise(total, seed=42),
        "uniform": signals.rand_noise(total, seed=43),
    }


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("window", [0, 1, 2])
def test_complex_bins_within_tolerance(built_lib, oracle_mod, n, window):
    from kopek_b200 import StftPlan, OUT_COMPLEX
    frames = 37
    with StftPlan(n, n, window, OUT_COMPLEX) as plan:
        for name, x in _signals(n, frames).items():
            got = _run(plan, x)[:, :plan.bins].astype(np.complex128)
            ref64 = oracle_mod.stft_complex(x, n, n, window, precision=64)
            ref32 = oracle_mod.stft_complex(x, n, n, window, precision=32)
            assert got.shape == ref64.shape == (frames, n // 2 + 1)
            e64, e32 = frame_rel_err(got, ref64), frame_rel_err(got, ref32)
            assert e64 <= tol(n), (name, e64)
            assert e32 <= tol(n), (name, e32)
            assert not peaks_match(np.abs(got), np.abs(ref64), n), name


@pytest.mark.parametrize("n", SIZES)
def test_magnitude_power_db(built_lib, oracle_mod, n):
    from kopek_b200 import StftPlan, OUT_MAG, OUT_POWER, OUT_DB, WINDOW_HANN
    frames = 19
    x = _signals(n, frames)["sine+noise"]
    ref = oracle_mod.stft_complex(x, n, n, WINDOW_HANN)
    mag_ref = np.abs(ref)
    peak = mag_ref.max(axis=1, keepdims=True)
    with StftPlan(n, n, WINDOW_HANN, OUT_MAG) as plan:
        mag = _run(plan, x)[:, :plan.bins]
        assert (np.abs(mag - mag_ref) / peak).max() <= tol(n)
        assert not peaks_match(mag, mag_ref, n)
    with StftPlan(n, n, WINDOW_HANN, OUT_POWER) as plan:
        pw = _run(plan, x)[:, :plan.bins]
        assert (np.abs(pw - mag_ref ** 2) / peak ** 2).max() <= 2 * tol(n)
    with StftPlan(n, n, WINDOW_HANN, OUT_DB) as plan:
        db = _run(plan, x)[:, :plan.bins]
        db_ref = 20 * np.log10(mag_ref)
        peak_db = 20 * np.log10(peak)
        # dB compared in the linear domain at the parity tolerance ...
        lin_err = np.abs(10 ** (db.astype(np.float64) / 20) - mag_ref) / peak
        assert lin_err.max() <= tol(n)
        # ... and in dB for bins within 40 dB of the frame's peak (bins below -120 dB are excluded anyway)
        strong = db_ref > peak_db - 40.0
        assert np.abs(db - db_ref)[strong].max() < 0.05


@pytest.mark.parametrize("n,hop", [(256, 64), (1024, 256), (2048, 512), (4096, 1024), (512, 130), (1024, 333),
                                   (256, 1), (1024, 1500), (2048, 514), (4096, 1026), (2048, 777), (4096, 5001),
                                   (4096, 2), (512, 4)])
def test_hop_addressing(built_lib, oracle_mod, n, hop):
    """STFT framing is an addressing mode of the load stage; odd hops take the scalar-load path of the generic kernel,
    hops that are 2 mod 4 the 64-bit loads of the 2048/4096-point team kernel."""
    from kopek_b200 import StftPlan, OUT_MAG, WINDOW_HANN, signals
    x = signals.chirp(n + hop * 45 + 17, 48000.0, 50.0, 15000.0)
    with StftPlan(n, hop, WINDOW_HANN, OUT_MAG) as plan:
        mag = _run(plan, x)[:, :plan.bins]
    ref = oracle_mod.stft_mag(x, n, hop, WINDOW_HANN)
    assert mag.shape == ref.shape == (1 + (x.size - n) // hop, n // 2 + 1)
    assert (np.abs(mag - ref) / ref.max(axis=1, keepdims=True)).max() <= tol(n)


def test_full_bins_pitch_and_edges(built_lib, oracle_mod):
    import torch
    from kopek_b200 import StftPlan, OUT_COMPLEX, OUT_MAG, BINS_FULL, WINDOW_RECT, WINDOW_HAMMING, signals
    n = 1024
    x = signals.white_noise(n * 9, seed=7)
    # all N bins, as the reference's callers iterate them (graph/utils.rs:50)
    with StftPlan(n, n, WINDOW_RECT, OUT_COMPLEX, BINS_FULL) as plan:
        assert plan.bins == n
        got = _run(plan, x).astype(np.complex128)
    ref = oracle_mod.stft_complex(x, n, n, 0, bins=n)
    assert frame_rel_err(got, ref) <= tol(n)
    # magnitude variant B over all bins (graph/utils.rs:53): sqrt(re^2 + im^2) in f32
    with StftPlan(n, n, WINDOW_RECT, OUT_MAG, BINS_FULL) as plan:
        mag = _run(plan, x)
    mag_b = np.stack([oracle_mod.mag_sqrt_f32(r) for r in ref])
    assert (np.abs(mag - mag_b) / mag_b.max(axis=1, keepdims=True)).max() <= tol(n)
    # padded pitch: rows start at multiples of pitch, padding untouched
    with StftPlan(n, n, WINDOW_HAMMING, OUT_MAG, out_pitch=520) as plan:
        assert plan.pitch == 520 and plan.bins == 513
        raw = _run(plan, x)
        assert np.isnan(raw[:, 513:]).all() and not np.isnan(raw[:, :513]).any()
        ref_m = oracle_mod.stft_mag(x, n, n, 2)
        assert (np.abs(raw[:, :513] - ref_m) / ref_m.max(axis=1, keepdims=True)).max() <= tol(n)
    # fewer samples than one frame -> zero frames, no launch; exactly one frame
    with StftPlan(n, n, WINDOW_RECT, OUT_MAG) as plan:
        assert plan.frames(n - 1) == 0
        d = torch.zeros(n - 1, device="cuda")
        o = torch.zeros(513, device="cuda")
        assert plan.exec_ptr(d.data_ptr(), n - 1, o.data_ptr()) == 0 and plan.last_launches == 0
        one = _run(plan, x[:n])
        assert one.shape[0] == 1
        # base pointer only 4-byte aligned (odd sample offset) -> scalar-load path
        dx = torch.from_numpy(x).cuda()
        o2 = torch.empty((3, 513), device="cuda")
        assert plan.exec_ptr(dx.data_ptr() + 4, 3 * n, o2.data_ptr()) == 3
        torch.cuda.synchronize()
        ref_o = oracle_mod.stft_mag(x[1:1 + 3 * n], n, n, 0)
        assert (np.abs(o2.cpu().numpy() - ref_o) / ref_o.max(axis=1, keepdims=True)).max() <= tol(n)


@pytest.mark.parametrize("n", SIZES)
def test_custom_window_every_size_and_output(built_lib, oracle_mod, n):
    """KOPEK_WINDOW_CUSTOM: the table-window variant of every kernel (for 2048/4096 the only user of that variant,
    since Hann and Hamming are evaluated on the fly there), on more frames than there are resident teams."""
    from kopek_b200 import StftPlan, OUT_MAG, OUT_COMPLEX, OUT_POWER, OUT_DB, WINDOW_CUSTOM
    rng = np.random.default_rng(n)
    frames = 2111
    w = rng.random(n).astype(np.float32)
    x = rng.standard_normal(n * frames).astype(np.float32)
    xw = (x.reshape(frames, n) * w).astype(np.float32).astype(np.float64)
    ref = np.fft.rfft(xw, axis=1)
    # the oracle (restated recursion of src/fft.rs) on the first and last frames; numpy carries the bulk
    for f in (0, 1, frames - 1):
        o = oracle_mod.fft(xw[f].astype(np.complex128))[: n // 2 + 1]
        assert np.abs(o - ref[f]).max() <= 1e-12 * np.abs(o).max()
    with StftPlan(n, n, WINDOW_CUSTOM, OUT_COMPLEX, custom_window=w) as plan:
        got = _run(plan, x).astype(np.complex128)
    assert frame_rel_err(got, ref) <= tol(n)
    top = np.abs(ref).max(axis=1, keepdims=True)
    with StftPlan(n, n, WINDOW_CUSTOM, OUT_MAG, custom_window=w) as plan:
        assert (np.abs(_run(plan, x) - np.abs(ref)) / top).max() <= tol(n)
    with StftPlan(n, n, WINDOW_CUSTOM, OUT_POWER, custom_window=w) as plan:
        assert (np.abs(_run(plan, x) - np.abs(ref) ** 2) / top ** 2).max() <= 2 * tol(n)
    with StftPlan(n, n, WINDOW_CUSTOM, OUT_DB, custom_window=w) as plan:
        assert (np.abs(10 ** (_run(plan, x).astype(np.float64) / 20) - np.abs(ref)) / top).max() <= tol(n)


def test_custom_window_and_analytic_answers(built_lib, oracle_mod):
    from kopek_b200 import StftPlan, OUT_MAG, OUT_COMPLEX, WINDOW_CUSTOM, WINDOW_RECT
    n = 512
    rng = np.random.default_rng(2)
    w = rng.random(n).astype(np.float32)
    x = rng.standard_normal(n * 5).astype(np.float32)
    with StftPlan(n, n, WINDOW_CUSTOM, OUT_COMPLEX, custom_window=w) as plan:
        got = _run(plan, x).astype(np.complex128)
    ref = np.fft.rfft((x.reshape(5, n) * w).astype(np.float32).astype(np.float64), axis=1)
    assert frame_rel_err(got, ref) <= tol(n)
    # impulse -> flat spectrum of ones; DC -> bin 0 = N
    for n in SIZES:
        imp = np.zeros(2 * n, dtype=np.float32); imp[0] = 1; imp[n:] = 1
        with StftPlan(n, n, WINDOW_RECT, OUT_MAG) as plan:
            m = _run(plan, imp)
        np.testing.assert_allclose(m[0], np.ones(n // 2 + 1), atol=1e-6)
        assert abs(m[1, 0] - n) <= 1e-4 * n and np.abs(m[1, 1:]).max() <= tol(n) * n


def test_golden_fixtures_through_abi(built_lib):
    from kopek_b200 import StftPlan, OUT_COMPLEX, OUT_DB, WINDOW_HANN, WINDOW_RECT
    g = np.load(os.path.join(GOLDEN, "sine440_1024.npz"))
    with StftPlan(1024, 1024, WINDOW_RECT, OUT_COMPLEX) as plan:
        got = _run(plan, g["samples"]).astype(np.complex128)[0]
    assert frame_rel_err(got, g["spectrum"][:513]) <= tol(1024)
    assert int(np.abs(got).argmax()) == int(g["peak_bin"]) == 10
    g = np.load(os.path.join(GOLDEN, "noise256.npz"))
    with StftPlan(256, 256, WINDOW_RECT, OUT_COMPLEX) as plan:
        got = _run(plan, g["samples"]).astype(np.complex128)
    assert frame_rel_err(got, g["spectrum"]) <= tol(256)
    g = np.load(os.path.join(GOLDEN, "chirp_hann2048.npz"))
    with StftPlan(2048, 512, WINDOW_HANN, OUT_DB) as plan:
        for fr, ref in zip(g["frames"], g["spectrum"]):
            db = _run(plan, fr)[0]
            mag_ref = np.abs(ref)
            assert (np.abs(10 ** (db / 20.0) - mag_ref) / mag_ref.max()).max() <= tol(2048)
            assert int(db.argmax()) == int(mag_ref.argmax())


def test_host_buffer_call_matches_device_call(built_lib, oracle_mod):
    """kopek_stft_exec_host (chunked copies on internal streams) == device-resident result, bit for bit."""
    from kopek_b200 import StftPlan, OUT_MAG, WINDOW_HANN, signals
    n, hop = 1024, 256
    x = signals.chirp(3_000_000, 48000.0, 20.0, 20000.0)
    with StftPlan(n, hop, WINDOW_HANN, OUT_MAG) as plan:
        a = _run(plan, x)
        b = plan.exec_host(x)
        assert plan.last_launches >= 1
    np.testing.assert_array_equal(a, b)
    ref = oracle_mod.stft_mag(x[: n + 99 * hop], n, hop, WINDOW_HANN)
    assert (np.abs(b[:100] - ref) / ref.max(axis=1, keepdims=True)).max() <= tol(n)


@pytest.mark.parametrize("n", SIZES)
def test_full_size_properties(built_lib, n):
    """BASELINE-size batches (2^28 samples per buffer: 2^18 frames of 1024, 2^20 of 256 = configs[2], ...): Parseval,
    linearity and a slice against torch.fft, on device -- every persistent kernel runs many frames per team here."""
    import torch
    from kopek_b200 import StftPlan, OUT_POWER, OUT_COMPLEX, WINDOW_RECT
    frames = (1 << 28) // n
    m = n // 2
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(frames * n, device="cuda", generator=g)
    y = torch.randn(frames * n, device="cuda", generator=g)
    st = torch.cuda.current_stream().cuda_stream
    with StftPlan(n, n, WINDOW_RECT, OUT_POWER) as plan:
        p = torch.empty((frames, m + 1), device="cuda")
        plan.exec_ptr(x.data_ptr(), x.numel(), p.data_ptr(), st)
        # Parseval for real input: sum x^2 = (P0 + P_M + 2 sum_{0<k<M} P_k) / N
        lhs = (x.view(frames, n).double() ** 2).sum(1)
        rhs = (p[:, 0].double() + p[:, m].double() + 2 * p[:, 1:m].double().sum(1)) / n
        assert ((lhs - rhs).abs() / lhs).max().item() < 1e-5
        del p, lhs, rhs
    with StftPlan(n, n, WINDOW_RECT, OUT_COMPLEX) as plan:
        fx = torch.empty((frames, m + 1), device="cuda", dtype=torch.complex64)
        fy = torch.empty_like(fx); fz = torch.empty_like(fx)
        z = 2.0 * x - 0.5 * y
        plan.exec_ptr(x.data_ptr(), x.numel(), fx.data_ptr(), st)
        plan.exec_ptr(y.data_ptr(), y.numel(), fy.data_ptr(), st)
        plan.exec_ptr(z.data_ptr(), z.numel(), fz.data_ptr(), st)
        torch.cuda.synchronize()
        lin = (fz - (2.0 * fx - 0.5 * fy)).abs().amax(1) / fz.abs().amax(1)
        assert lin.max().item() < 2e-5
        # checksum of checksums vs torch.fft on a slice (library as a second opinion, not the oracle)
        ref = torch.fft.rfft(x.view(frames, n)[-4096:].double(), dim=1)
        err = (fx[-4096:].to(torch.complex128) - ref).abs().amax(1) / ref.abs().amax(1)
        assert err.max().item() <= tol(n)


@pytest.mark.parametrize("window,output", [(1, 1), (0, 3), (2, 2)])
def test_bulk_store_epilogue_is_bit_identical(built_lib, window, output):
    """kopek_plan_set_bulk_store: one 2048-byte TMA store per frame instead of 32-bit stores -- same bits, padding untouched."""
    import torch
    from kopek_b200 import StftPlan, KopekError
    n, frames = 1024, 3001
    x = torch.randn(frames * n, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ref_plan = StftPlan(n, n, window, output, out_pitch=516)
    bulk_plan = StftPlan(n, n, window, output, out_pitch=516, bulk_store=True)
    if True:
        a = torch.full((frames, 516), float("nan"), device="cuda")
        b = torch.full((frames, 516), float("nan"), device="cuda")
        ref_plan.exec_ptr(x.data_ptr(), x.numel(), a.data_ptr(), st)
        bulk_plan.exec_ptr(x.data_ptr(), x.numel(), b.data_ptr(), st)
        torch.cuda.synchronize()
        assert torch.equal(a[:, :513], b[:, :513])
        assert bool(torch.isnan(b[:, 513:]).all()) and not bool(torch.isnan(b[:, :513]).any())
        # unaligned output pointer -> silently the direct stores, still correct
        c = torch.full((frames * 516 + 1,), float("nan"), device="cuda")
        bulk_plan.exec_ptr(x.data_ptr(), x.numel(), c.data_ptr() + 4, st)
        torch.cuda.synchronize()
        assert torch.equal(c[1:].view(frames, 516)[:, :513], a[:, :513])
    ref_plan.close()
    bulk_plan.close()
    with pytest.raises(KopekError):
        StftPlan(n, n, window, output, bulk_store=True)            # pitch 513: rows not 16-byte aligned
    with pytest.raises(KopekError):
        StftPlan(512, 512, window, output, out_pitch=260, bulk_store=True)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("window", [1, 2])
def test_window_is_accurate_sample_by_sample(built_lib, oracle_mod, n, window):
    """One impulse per frame: |X[k]| = w[p] for every bin, so each frame checks ONE window coefficient in isolation
    (relative to itself, not to the frame of a broadband signal).  The 2048/4096-point kernels evaluate Hann and
    Hamming as a cosine sum on the fly (fft_team.cuh); this pins that evaluation to the oracle's f32 table at the
    frame edges, where the window is ~1e-7, at the seams of the per-lane segments, and at random positions."""
    from kopek_b200 import StftPlan, OUT_MAG
    rng = np.random.default_rng(n + window)
    seg = n // 16
    pos = sorted({0, 1, 2, 3, 5, seg - 2, seg - 1, seg, seg + 1, 2 * seg - 1, 2 * seg, n // 2 - 1, n // 2, n // 2 + 1,
                  n - 2 * seg - 1, n - 2 * seg, n - seg - 1, n - seg, n - seg + 1, n - 3, n - 2, n - 1}
                 | set(int(v) for v in rng.integers(0, n, 64)))
    x = np.zeros((len(pos), n), dtype=np.float32)
    for f, p in enumerate(pos):
        x[f, p] = 1.0 if f % 2 == 0 else -0.75
    x = x.reshape(-1)
    w = oracle_mod.window(window, n).astype(np.float64)
    with StftPlan(n, n, window, OUT_MAG) as plan:
        got = _run(plan, x)[:, :plan.bins].astype(np.float64)
    for f, p in enumerate(pos):
        ref = abs(x[f * n + p]) * w[p]
        if ref == 0.0:
            assert np.abs(got[f]).max() <= 1e-9, p
        else:
            assert np.abs(got[f] - ref).max() <= tol(n) * ref, (p, got[f, :3], ref)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("output,window", [(0, 0), (1, 1), (3, 2), (2, 0)])
def test_full_rows_every_size(built_lib, oracle_mod, n, output, window):
    """All N bins per row, as the reference's callers iterate them (examples/graph/utils.rs:47-63): the lane-group
    kernels store the conjugate mirror beside the computed half.  Held to the oracle's full-length spectrum, to exact
    mirror symmetry, and to the guard values around every row (pitch > N)."""
    from kopek_b200 import BINS_FULL, StftPlan, signals
    hop = n // 2 + 4
    x = signals.white_noise(n + hop * 37, seed=n + output) + 0.5
    with StftPlan(n, hop, window, output, BINS_FULL, out_pitch=n + 8) as plan:
        assert plan.bins == n and plan.pitch == n + 8
        raw = _run(plan, x)
    assert np.isnan(raw[:, n:].real).all(), "wrote into the row padding"
    got = raw[:, :n]
    ref = oracle_mod.stft_complex(x, n, hop, window, bins=n)
    if output == 0:
        assert frame_rel_err(got.astype(np.complex128), ref) <= tol(n)
        assert np.array_equal(got[:, 1:n // 2], np.conj(got[:, :n // 2:-1]))           # X[N - k] == conj X[k], bit for bit
    else:
        lin = np.abs(ref)
        g = got.astype(np.float64)
        g = 10.0 ** (g / 20.0) if output == 3 else np.sqrt(g) if output == 2 else g
        assert (np.abs(g - lin) / lin.max(axis=1, keepdims=True)).max() <= tol(n)
        assert np.array_equal(got[:, 1:n // 2], got[:, :n // 2:-1])


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("hop_kind", ["odd", "two_mod_four"])
def test_unaligned_frames_every_size(built_lib, oracle_mod, n, hop_kind, monkeypatch):
    """Odd hops and base pointers that are only 4-byte aligned run on the lane-group kernels with 32-bit loads: same
    values as the vector-load instantiation (bit for bit, on an aligned copy of the same samples) and within tolerance
    of the oracle and of the generic kernel."""
    import torch
    from kopek_b200 import StftPlan, signals
    hop = n // 4 + (3 if hop_kind == "odd" else 2)
    frames = 53
    x = signals.chirp((frames - 1) * hop + n + 3, 48000.0, 30.0, 21000.0) + 0.01 * signals.white_noise((frames - 1) * hop + n + 3, 3)
    dx = torch.from_numpy(x).cuda()
    st = torch.cuda.current_stream().cuda_stream
    with StftPlan(n, hop, 1, 1) as plan:
        out = torch.full((frames, plan.pitch), float("nan"), device="cuda")
        # base shifted by one sample: 4-byte aligned only
        assert plan.exec_ptr(dx.data_ptr() + 4, (frames - 1) * hop + n, out.data_ptr(), st) == frames
        torch.cuda.synchronize()
        got = out.cpu().numpy()
    ref = oracle_mod.stft_mag(x[1:], n, hop, 1)[:frames]
    assert (np.abs(got - ref) / ref.max(axis=1, keepdims=True)).max() <= tol(n)
    # frame by frame on an aligned copy (hop = n: the vector-load instantiation): identical bits
    fr = np.stack([x[1 + f * hop: 1 + f * hop + n] for f in range(frames)]).reshape(-1)
    with StftPlan(n, n, 1, 1) as vplan:
        same = _run(vplan, fr)
    assert np.array_equal(got.view(np.uint32), same.view(np.uint32))
    # the generic kernel (cross-check build of the same plan)
    monkeypatch.setenv("KOPEK_K1W_VARIANT", "-1")
    with StftPlan(n, hop, 1, 1) as gplan:
        gen = torch.full((frames, gplan.pitch), float("nan"), device="cuda")
        gplan.exec_ptr(dx.data_ptr() + 4, (frames - 1) * hop + n, gen.data_ptr(), st)
        torch.cuda.synchronize()
    assert (np.abs(gen.cpu().numpy() - ref) / ref.max(axis=1, keepdims=True)).max() <= tol(n)


@pytest.mark.parametrize("n", [2048, 4096])
def test_custom_window_full_rows_and_odd_hop_on_the_team_kernels(built_lib, oracle_mod, n):
    """Round 2: custom windows run from a shared-memory table in the team kernels, so they too get FULL rows and
    unaligned frames (the last corner that fell back to the generic kernel)."""
    from kopek_b200 import BINS_FULL, StftPlan, WINDOW_CUSTOM, signals
    hop = n // 2 + 1
    x = signals.white_noise(n + hop * 21, seed=n) + 0.25
    w = np.blackman(n).astype(np.float32)
    with StftPlan(n, hop, WINDOW_CUSTOM, 1, BINS_FULL, custom_window=w) as plan:
        got = _run(plan, x)[:, :n]
    frames = 1 + (x.size - n) // hop
    ref = np.stack([np.abs(oracle_mod.fft((x[f * hop: f * hop + n] * w).astype(np.float32).astype(np.complex128)))
                    for f in range(frames)])
    assert got.shape == ref.shape
    assert (np.abs(got - ref) / ref.max(axis=1, keepdims=True)).max() <= tol(n)
    assert np.array_equal(got[:, 1:n // 2], got[:, :n // 2:-1])


@pytest.mark.parametrize("n_in,n", [(441, 512), (1000, 1024), (480, 512), (3000, 4096)])
def test_zero_padded_callback_frames_through_a_custom_window(built_lib, oracle_mod, n_in, n):
    """The graph example hands fft() whatever length the audio callback delivered (examples/graph/player.rs:48-61), and
    fft() pads it with zeros to the next power of two (src/fft.rs:31-36).  Many such frames at once: hop = n_in, frame
    size n, and a custom window that is 1 on [0, n_in) and 0 on the tail -- the tail samples (the next frame's) are
    multiplied by zero, which IS the zero padding.  Checked against fft() of every padded frame."""
    from kopek_b200 import StftPlan, WINDOW_CUSTOM, signals
    frames = 29
    x = signals.chirp(frames * n_in + (n - n_in), 44100.0, 100.0, 9000.0)
    w = np.zeros(n, dtype=np.float32)
    w[:n_in] = 1.0
    with StftPlan(n, n_in, WINDOW_CUSTOM, 1, custom_window=w) as plan:
        got = _run(plan, x)[:, :plan.bins]
    assert got.shape == (frames, n // 2 + 1)
    ref = np.stack([oracle_mod.mag_norm(oracle_mod.fft(oracle_mod.marshal(x[f * n_in:(f + 1) * n_in])))[: n // 2 + 1]
                    for f in range(frames)])
    assert (np.abs(got - ref) / ref.max(axis=1, keepdims=True)).max() <= tol(n)
    assert np.array_equal(got.argmax(axis=1), ref.argmax(axis=1))
