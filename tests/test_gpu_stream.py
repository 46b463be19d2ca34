"""GPU: the live sliding-window analyser (kopek_stream_*) against the oracle -- the beep view's
VecDeque of the latest 1024 samples (examples/beep/view.rs:44,57-60) transformed every UI frame (:140-158)."""
import numpy as np
import pytest

from util import tol

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,window,scale", [(1024, 0, 1.0), (1024, 0, 1.0 / 32767.0), (1024, 1, 1.0), (256, 2, 0.5),
                                            (4096, 1, 1.0)])
def test_live_spectrum_follows_the_deque(built_lib, oracle_mod, n, window, scale):
    from kopek_b200 import LiveSpectrum, OUT_MAG
    rng = np.random.default_rng(n + window)
    live = LiveSpectrum(n, window, OUT_MAG, input_scale=scale)
    deque = np.zeros(n, dtype=np.float32)                       # VecDeque initialised with n zeros
    w = oracle_mod.window(window, n)
    # generator pushes arbitrary chunk sizes between UI frames (ring of 100000 in beep/view.rs:26-31)
    for count in [0, 1, 3, 4, 17, 1024 if n >= 1024 else 100, 511, n, n + 5, 3 * n + 1, 2, 64, 1000, 7]:
        new = oscil = (0.5 * np.sin(2 * np.pi * 440.0 * (np.arange(count) + rng.integers(0, 1000)) / 48000.0)
                       + 0.1 * rng.standard_normal(count)).astype(np.float32)
        deque = np.concatenate([deque, new])[-n:]
        got = live.push(new)
        # reference: fft_input = deque.map(|s| s as f64 * scale) (beep/view.rs:140-144), fft, norm()
        xin = (deque * w).astype(np.float32).astype(np.float64) * np.float64(scale)
        ref = oracle_mod.mag_norm(oracle_mod.fft(xin.astype(np.complex128)))[: n // 2 + 1]
        assert got.shape == ref.shape
        den = max(ref.max(), 1e-30)
        assert np.abs(got - ref).max() / den <= tol(n), count
        if ref.max() > 0:
            assert int(got.argmax()) == int(ref.argmax()) or abs(ref[got.argmax()] - ref.max()) <= 2 * tol(n) * ref.max()
    live.close()


def test_live_spectrum_complex_and_errors(built_lib, oracle_mod):
    from kopek_b200 import LiveSpectrum, OUT_COMPLEX, KopekError
    live = LiveSpectrum(512, 0, OUT_COMPLEX)
    x = np.random.default_rng(1).standard_normal(512).astype(np.float32)
    got = live.push(x).astype(np.complex128)
    ref = oracle_mod.fft(x.astype(np.complex128))[:257]
    assert np.abs(got - ref).max() / np.abs(ref).max() <= tol(512)
    with pytest.raises(KopekError):
        LiveSpectrum(1000)                       # not a supported frame size
    with pytest.raises(KopekError):
        LiveSpectrum(1024, window=3)             # custom windows are not a stream option


def test_live_bands_with_unaligned_push_sizes(built_lib, oracle_mod):
    """441-sample pushes (one UI frame of 44.1 kHz audio at 100 Hz) make the cumulative count odd; the
    window handed to the kernel stays aligned, so the band epilogue keeps working and every push runs the same kernel."""
    from kopek_b200 import LiveSpectrum, OUT_BANDS_LOG, OUT_MAG
    rng = np.random.default_rng(9)
    bands, mags = LiveSpectrum(1024, 0, OUT_BANDS_LOG), LiveSpectrum(1024, 0, OUT_MAG)
    for count in (441, 441, 1, 441, 3, 1024, 441):
        new = rng.standard_normal(count).astype(np.float32)
        row, y = bands.push(new), mags.push(new)
        want = oracle_mod.bands_log(y.astype(np.float32)[:512])
        assert row.shape == (10,)
        assert np.abs(row[:8] - want).max() <= 1e-4 * want.max()
    bands.close()
    mags.close()
