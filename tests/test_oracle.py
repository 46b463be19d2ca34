"""CPU: pin the oracle (restatement of src/fft.rs) against independent implementations."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from util import frame_rel_err


@pytest.mark.parametrize("n", [1, 2, 4, 8, 16, 64, 256, 1024, 4096])
def test_oracle_vs_naive_dft_and_numpy(oracle_mod, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    y = oracle_mod.fft(x)
    assert y.shape == (n,)
    assert frame_rel_err(y, np.fft.fft(x)) < 4e-15
    if n <= 1024:
        assert frame_rel_err(y, oracle_mod.dft_naive(x)) < 4e-15


@pytest.mark.parametrize("n_in", [0, 1, 3, 5, 100, 1000, 1025])
def test_oracle_zero_pads_to_next_pow2(oracle_mod, n_in):
    # src/fft.rs:31-36; len 0 -> n = 1 -> [0+0i]
    rng = np.random.default_rng(n_in + 1)
    x = rng.standard_normal(n_in) + 1j * rng.standard_normal(n_in)
    y = oracle_mod.fft(x)
    n = oracle_mod.next_pow2(n_in)
    assert n >= 1 and (n & (n - 1)) == 0 and n >= n_in and (n_in <= 1 or n < 2 * n_in)
    assert y.shape == (n,)
    if n_in == 0:
        assert y[0] == 0
    else:
        assert frame_rel_err(y, np.fft.fft(x, n)) < 4e-15


def test_oracle_f32_instantiation_within_tolerance(oracle_mod):
    rng = np.random.default_rng(3)
    for n in (256, 1024, 4096):
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        y32 = oracle_mod.fft(x, np.complex64)
        assert y32.dtype == np.complex64
        assert frame_rel_err(y32.astype(np.complex128), np.fft.fft(x.astype(np.complex128))) < 1e-5 * np.log2(n)


def test_golden_sine440(oracle_mod):
    g = np.load(os.path.join(GOLDEN, "sine440_1024.npz"))
    s = oracle_mod.osc_sine(44100.0, 440.0, 1024)
    np.testing.assert_array_equal(s, g["samples"])           # oscillator restatement is deterministic
    X = oracle_mod.fft(oracle_mod.marshal(s))
    assert frame_rel_err(X, g["spectrum"]) < 4e-15
    mag = oracle_mod.mag_norm(X)
    assert int(mag[:513].argmax()) == int(g["peak_bin"]) == 10   # 440*1024/44100 = 10.2
    assert abs(mag.max() - 471.9957) < 1e-3


def test_golden_ragged_and_noise(oracle_mod):
    g = np.load(os.path.join(GOLDEN, "ragged1000.npz"))
    assert frame_rel_err(oracle_mod.fft(g["input"]), g["spectrum"]) < 4e-15
    g = np.load(os.path.join(GOLDEN, "noise256.npz"))
    got = oracle_mod.stft_complex(g["samples"], 256, 256)
    assert got.shape == (4, 129)
    assert frame_rel_err(got, g["spectrum"]) < 4e-15


def test_golden_chirp_hann(oracle_mod):
    g = np.load(os.path.join(GOLDEN, "chirp_hann2048.npz"))
    for fr, ref in zip(g["frames"], g["spectrum"]):
        got = oracle_mod.stft_complex(fr, 2048, 512, oracle_mod.WIN_HANN)
        assert got.shape == (1, 1025)
        assert frame_rel_err(got, ref[None, :]) < 4e-15


def test_show_format(oracle_mod):
    want = open(os.path.join(GOLDEN, "show.txt")).read()
    got = oracle_mod.show("label", [1 + 2j, -0.5 - 0.25j, 0j, complex(-0.0, -3.14159)])
    assert got == want
    assert oracle_mod.show("x", [complex(float("nan"), float("inf"))]) == "x\nNaN+infi\n"


def test_oscillator_restatement(oracle_mod):
    # closed form sin(2 pi f t / sr) within the f32 accumulator's drift (src/oscillator.rs:47-53)
    s = oracle_mod.osc_sine(48000.0, 1000.0, 4800)
    t = np.arange(4800) / 48000.0
    assert np.abs(s - np.sin(2 * np.pi * 1000.0 * t)).max() < 2e-3
    assert s[0] == 0.0
    for wave in (1, 2, 3, 4):
        w = oracle_mod.osc_wave(wave, 48000.0, 440.0, 1000)
        assert np.all(np.abs(w) <= 1.0 + 1e-6)


def test_magnitude_variants_and_bands(oracle_mod):
    rng = np.random.default_rng(9)
    X = rng.standard_normal(1024) + 1j * rng.standard_normal(1024)
    a = oracle_mod.mag_norm(X)           # beep/view.rs:155
    b = oracle_mod.mag_sqrt_f32(X)       # graph/utils.rs:53
    np.testing.assert_allclose(a, np.abs(X), rtol=1e-15)
    np.testing.assert_allclose(b, np.abs(X), rtol=3e-7)
    y = (b * 100.0).astype(np.float32)
    low = oracle_mod.bands_low(y)
    np.testing.assert_allclose(low, [y[20 + 3 * i:23 + 3 * i].mean() for i in range(8)], rtol=1e-6)
    logb = oracle_mod.bands_log(y)
    edges = np.cumsum([0, 4, 4, 8, 16, 32, 64, 128, 256])
    np.testing.assert_allclose(logb, [y[edges[i]:edges[i + 1]].mean() for i in range(8)], rtol=1e-5)
    idx, mx = oracle_mod.peak_band(low)
    assert idx == int(np.argmax(low)) and mx == pytest.approx(float(low.max()))


def test_stft_composition(oracle_mod):
    rng = np.random.default_rng(11)
    x = rng.standard_normal(6000).astype(np.float32)
    for n, hop, win in [(256, 256, 0), (512, 128, 1), (1024, 1000, 2)]:
        w = oracle_mod.window(win, n)
        frames = oracle_mod.stft_frames(x.size, n, hop)
        assert frames == 1 + (x.size - n) // hop
        ref = np.stack([np.fft.rfft((x[f * hop:f * hop + n] * w).astype(np.float32).astype(np.float64))
                        for f in range(frames)])
        for threads in (1, 3):
            got = oracle_mod.stft_complex(x, n, hop, win, threads=threads)
            assert frame_rel_err(got, ref) < 4e-15
        mag = oracle_mod.stft_mag(x, n, hop, win, output=oracle_mod.OUT_MAG)
        np.testing.assert_allclose(mag, np.abs(ref), rtol=2e-7, atol=1e-12)
        db = oracle_mod.stft_mag(x, n, hop, win, output=oracle_mod.OUT_DB)
        np.testing.assert_allclose(db, 20 * np.log10(np.abs(ref)), rtol=1e-6, atol=1e-5)
    assert oracle_mod.stft_frames(100, 256, 64) == 0
    wh = oracle_mod.window(oracle_mod.WIN_HANN, 8)
    np.testing.assert_allclose(wh, 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(8) / 8), atol=1e-7)
