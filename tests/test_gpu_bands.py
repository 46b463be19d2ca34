"""GPU: fused band-energy + peak-pick epilogue against the oracle's restatement of
examples/graph/utils.rs:47-116, examples/pong/utils.rs:36-114 and examples/pong/main.rs:192-213."""
import numpy as np
import pytest

from util import tol

pytestmark = pytest.mark.gpu


def _run(plan, x):
    import torch
    d_in = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda()
    frames = plan.frames(d_in.numel())
    out = torch.full((frames, plan.pitch), float("nan"), device="cuda")
    assert plan.exec_ptr(d_in.data_ptr(), d_in.numel(), out.data_ptr(), torch.cuda.current_stream().cuda_stream) == frames
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _oracle_rows(oracle_mod, x, frames, kind, scale, div=1.0):
    rows = []
    for f in range(frames):
        # marshalling f32 -> f64 -> Complex{re,0} (pong/main.rs:192-195; div = 32767 in beep/graph), fft, magnitude B * scale
        X = oracle_mod.fft(oracle_mod.marshal(x[f * 1024:(f + 1) * 1024], div))
        y = (oracle_mod.mag_sqrt_f32(X) * np.float32(scale)).astype(np.float32)
        avg = oracle_mod.bands_low(y) if kind == "low" else oracle_mod.bands_log(y)
        idx, mx = oracle_mod.peak_band(avg)
        rows.append(np.concatenate([avg, [idx, mx]]))
    return np.array(rows, dtype=np.float64)


@pytest.mark.parametrize("kind,scale", [("low", 100.0), ("log", 10000.0)])
def test_band_epilogue_matches_reference_callers(built_lib, oracle_mod, kind, scale):
    from kopek_b200 import StftPlan, OUT_BANDS_LOG, OUT_BANDS_LOW, WINDOW_RECT, BAND_ROW, signals
    frames = 53
    out_mode = OUT_BANDS_LOW if kind == "low" else OUT_BANDS_LOG
    for name, x in {
        "tones": signals.sine_frames(frames, 1024, sample_rate=48000.0, noise=0.05, seed=3),
        "white": signals.white_noise(frames * 1024, seed=5),
        "chirp": signals.chirp(frames * 1024, 48000.0, 300.0, 3000.0),
    }.items():
        with StftPlan(1024, 1024, WINDOW_RECT, out_mode, band_scale=scale) as plan:
            assert plan.bins == BAND_ROW == 10 and plan.pitch == 10
            got = _run(plan, x).astype(np.float64)
        ref = _oracle_rows(oracle_mod, x, frames, kind, scale)
        # band means: tolerance relative to the frame's largest band mean (sum order differs from the
        # reference's sequential f32 sum, rounding ~1e-7)
        den = np.maximum(ref[:, :8].max(axis=1, keepdims=True), 1e-30)
        assert (np.abs(got[:, :8] - ref[:, :8]) / den).max() <= tol(1024), name
        # peak pick: same band unless the oracle's two best means are closer than the tolerance
        for f in range(frames):
            if got[f, 8] != ref[f, 8]:
                top2 = np.sort(ref[f, :8])[-2:]
                assert abs(top2[1] - top2[0]) <= 2 * tol(1024) * top2[1], (name, f)
            assert abs(got[f, 9] - ref[f, 9]) <= tol(1024) * max(ref[f, 9], 1e-30)


def test_band_epilogue_pong_threshold_and_silence(built_lib, oracle_mod):
    """pong/main.rs:200-213: silence -> max stays 0, index 0 -> factor 0; a tone in band 3 of the low set."""
    from kopek_b200 import StftPlan, OUT_BANDS_LOW, WINDOW_RECT
    sr = 44100.0
    tone_bin = 30                                    # band (30 - 20) // 3 = 3
    t = np.arange(1024) / sr
    x = np.concatenate([np.zeros(1024), np.sin(2 * np.pi * tone_bin * sr / 1024 * t)]).astype(np.float32)
    with StftPlan(1024, 1024, WINDOW_RECT, OUT_BANDS_LOW, band_scale=100.0) as plan:
        got = _run(plan, x)
    assert np.all(got[0, :8] == 0) and got[0, 8] == 0 and got[0, 9] == 0
    assert got[1, 8] == 3 and got[1, 9] > 50.0       # above the paddle threshold
    ref = _oracle_rows(oracle_mod, x, 2, "low", 100.0)
    assert ref[1, 8] == 3


def test_band_epilogue_argument_checks(built_lib):
    from kopek_b200 import StftPlan, OUT_BANDS_LOG, KopekError, BINS_FULL
    with pytest.raises(KopekError):
        StftPlan(512, 512, 0, OUT_BANDS_LOG)
    with pytest.raises(KopekError):
        StftPlan(1024, 1024, 0, OUT_BANDS_LOG, bins=BINS_FULL)


def test_band_epilogue_with_an_odd_hop(built_lib, oracle_mod):
    """Round 2: odd hops run on the same kernel (32-bit loads), so the band rows of overlapping, unaligned frames
    equal those of the same frames laid out back to back."""
    from kopek_b200 import StftPlan, OUT_BANDS_LOG, WINDOW_HANN, signals
    hop, frames = 1023, 40
    x = signals.white_noise((frames - 1) * hop + 1024, seed=11)
    with StftPlan(1024, hop, WINDOW_HANN, OUT_BANDS_LOG, band_scale=100.0) as plan:
        got = _run(plan, x)
    fr = np.stack([x[f * hop: f * hop + 1024] for f in range(frames)]).reshape(-1)
    with StftPlan(1024, 1024, WINDOW_HANN, OUT_BANDS_LOG, band_scale=100.0) as plan:
        ref = _run(plan, fr)
    assert got.shape == ref.shape == (frames, 10)
    assert np.array_equal(got, ref)
