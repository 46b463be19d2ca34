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


def test_reference_vectors_from_rust(oracle_mod):
    """The pin: outputs of the UNMODIFIED Rust reference (rust/golden, `cargo run` anywhere with a toolchain) on the
    committed inputs.  Bit equality with the oracle for fft() (src/fft.rs:6-41), text equality for show()'s bin format
    (src/fft.rs:43-52), the value of detect_bpm (src/lib.rs:33-38).  Until somebody has run the kit the files are
    absent and parity stays UNPINNED: the skip says so loudly."""
    import glob
    ins = sorted(glob.glob(os.path.join(GOLDEN, "rust_in_*.bin")))
    assert len(ins) >= 6, "the committed inputs of the kit are missing: python tests/golden/make_golden.py"
    outs = [p.replace("rust_in_", "rust_fft_") for p in ins]
    if not all(os.path.exists(p) for p in outs):
        pytest.skip("PARITY UNPINNED: tests/golden/rust_fft_*.bin absent -- run `cargo run --release --manifest-path "
                    "rust/golden/Cargo.toml -- tests/golden <sample.wav>` where a Rust toolchain exists")
    for pin, pout in zip(ins, outs):
        z = np.fromfile(pin, dtype="<c16")
        want = np.fromfile(pout, dtype="<c16")
        got = oracle_mod.fft(z)
        assert got.size == want.size == oracle_mod.next_pow2(z.size), pin
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), f"{pin}: oracle differs from the Rust reference"
    show = os.path.join(GOLDEN, "rust_show.txt")
    if os.path.exists(show):
        tricky = [1 + 2j, -0.5 - 0.25j, 0j, complex(-0.0, -3.14159), complex(float("nan"), float("-inf")),
                  complex(float("inf"), 0.00004), complex(123456.78905, -0.00005)]
        assert oracle_mod.show("label", tricky) == open(show).read()
    bpm = os.path.join(GOLDEN, "rust_detect_bpm.txt")
    wav = "/root/reference/assets/audio_samples/sample.wav"
    if os.path.exists(bpm) and os.path.exists(wav):
        import wave
        with wave.open(wav) as w:
            fr = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").reshape(-1, 2)
        assert oracle_mod.detect_bpm(fr, 44100.0)[0] == int(open(bpm).read().strip())


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


# ---- decoder-side callers: src/decoder.rs:39-71, examples/graph/player.rs:56-59 ---------------
def _click_track(seconds, bpm, seed, sr=44100):
    """Stereo i16 test signal: low-level noise with short loud bursts at `bpm`."""
    rng = np.random.default_rng(seed)
    n = int(seconds * sr) + 777
    x = (rng.standard_normal((n, 2)) * 300).astype(np.int16)
    period = int(sr * 60 / bpm)
    for s in range(0, n - 900, period):
        x[s:s + 900] += (rng.standard_normal((900, 2)) * 9000).astype(np.int16)
    return x


def _bpm_numpy(frames, sr=44100.0):
    """Independent restatement of detect_bpm with numpy f32 cumulative sums (sequential order)."""
    f = frames.astype(np.float32) / np.float32(32767)
    dur = int(np.float32(len(frames)) / np.float32(sr))
    pw = (f[:, 0] * f[:, 0] + f[:, 1] * f[:, 1]).astype(np.float32)
    beats, blocks, avgs = 0, [], []
    for i in range(dur):
        seg = pw[i * 44100:(i + 1) * 44100]
        avg = np.float32(np.cumsum(seg, dtype=np.float32)[-1]) * np.float32(np.float32(1024.0) / np.float32(44100.0))
        avgs.append(avg)
        for j in range(43):
            e = np.cumsum(seg[j * 1024:(j + 1) * 1024], dtype=np.float32)[-1]
            blocks.append(e)
            beats += bool(e > np.float32(5.5) * avg)
    return beats * (60 // dur), beats, np.array(blocks, np.float32), np.array(avgs, np.float32)


def test_detect_bpm_restatement(oracle_mod):
    for seconds, bpm, seed in ((3.2, 120, 0), (7.9, 90, 1), (1.0, 60, 2), (12.5, 174, 3)):
        x = _click_track(seconds, bpm, seed)
        got = oracle_mod.detect_bpm(x)
        ref = _bpm_numpy(x)
        assert got[0] == ref[0] and got[1] == ref[1]
        np.testing.assert_array_equal(got[2], ref[2])
        np.testing.assert_array_equal(got[3], ref[3])
    # where the Rust original panics: less than a second (60 / 0), or sample_rate < 44100 (slice out of range)
    assert oracle_mod.detect_bpm(_click_track(0.5, 120, 4)[:20000])[0] is None
    assert oracle_mod.detect_bpm(_click_track(2.0, 120, 5), sample_rate=22050.0)[0] is None


def test_detect_bpm_on_the_reference_asset(oracle_mod):
    """The reference's own test (src/lib.rs:33-38) expects 83 for assets/audio_samples/sample.wav, next to a
    duration test (src/lib.rs:27-31) that expects 2.0 s for 124443 frames: both are stale (124443 / 44100 = 2.82 s,
    and beats * (60 / 2) is a multiple of 30).  The restated algorithm and an independent numpy restatement agree on
    6 beats -> 180; the frame count and first frame of the asset (src/lib.rs:15-25) do match."""
    import wave
    path = "/root/reference/assets/audio_samples/sample.wav"
    if not os.path.exists(path):
        pytest.skip("reference assets are not on this machine")
    with wave.open(path) as w:
        assert (w.getnchannels(), w.getframerate(), w.getsampwidth(), w.getnframes()) == (2, 44100, 2, 124443)
        frames = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").reshape(-1, 2)
    assert frames[0].tolist() == [79, 79] and len(frames) == 124443          # src/lib.rs:15-25
    bpm, beats, block_e, avg_e = oracle_mod.detect_bpm(frames)
    ref = _bpm_numpy(frames)
    assert (bpm, beats) == (ref[0], ref[1]) == (180, 6)
    np.testing.assert_array_equal(block_e, ref[2])


def test_pcm16_marshalling(oracle_mod):
    rng = np.random.default_rng(9)
    pcm = rng.integers(-32768, 32768, size=(5000, 2), dtype=np.int16)
    pcm[:4, 0] = [32767, -32768, 0, 1]
    for ch in (0, 1):
        got = oracle_mod.pcm16_to_f32(pcm, ch)
        ref = (pcm[:, ch].astype(np.float64) / 32767.0).astype(np.float32)
        np.testing.assert_array_equal(got, ref)
    assert oracle_mod.pcm16_to_f32(pcm, 0)[0] == 1.0
    # the identity the fused load stage of fft_warp1024b.cuh relies on, for ALL 65536 inputs:
    # q = s r, q += fma(-q, 32767, s) r with r = RN(1/32767)  ==  (f32)((f64)s / 32767)   (and so is a plain f32 division)
    s_all = np.arange(-32768, 32768, dtype=np.int16)
    ref = oracle_mod.pcm16_to_f32(s_all)
    sf = s_all.astype(np.float32)
    r = np.float32(1.0) / np.float32(32767.0)
    q0 = (sf * r).astype(np.float32)
    e = (sf.astype(np.float64) - q0.astype(np.float64) * 32767.0).astype(np.float32)      # fma: one rounding
    q1 = (e.astype(np.float64) * np.float64(r) + q0.astype(np.float64)).astype(np.float32)
    np.testing.assert_array_equal(q1, ref)
    np.testing.assert_array_equal((sf / np.float32(32767.0)).astype(np.float32), ref)


def test_load_track_scaling(oracle_mod):
    """examples/graph/player.rs:81-102: volume_factor * frame[c] as f32, flattened -- an f32 product per value."""
    allv = np.arange(-32768, 32768, dtype=np.int16)
    got = oracle_mod.pcm16_scale_f32(allv, 0.00002)
    assert np.array_equal(got.view(np.uint32), (np.float32(0.00002) * allv.astype(np.float32)).view(np.uint32))
    assert got[32768] == 0.0 and got[-1] == np.float32(0.00002) * np.float32(32767)


def test_capture_downmix(oracle_mod):
    """examples/pong/audio.rs:57-70: the FIRST TWO channels are summed, the sum is divided by ALL channels."""
    f = np.array([[1.0, 3.0, 100.0, 100.0], [-0.0, -0.0, 5.0, 5.0]], dtype=np.float32)
    got = oracle_mod.f32_downmix(f, 4)
    assert got[0] == 1.0 and got[1] == 0.0 and not np.signbit(got[1])       # 0.0 + -0.0 = +0.0
    assert oracle_mod.f32_downmix(np.array([2.5, -1.5], dtype=np.float32), 1).tolist() == [2.5, -1.5]
