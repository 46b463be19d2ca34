"""GPU: on-device signal sources (kopek_synth_frames) against the oracle's restatement of src/oscillator.rs / src/noise.rs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gen(n_frames, n, sr, **kw):
    import torch
    from kopek_b200 import fft
    out = torch.full((n_frames, n), float("nan"), device="cuda")
    fft.synth_frames_device(out.data_ptr(), n_frames, n, sr, stream=torch.cuda.current_stream().cuda_stream, **kw)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def test_oscillator_waves_match_the_recurrence(built_lib, oracle_mod):
    from kopek_b200 import fft
    n, sr, frames = 1024, 48000.0, 40
    freqs = 20.0 + 20.0 * (np.arange(frames) % 1000)                      # workload M's frame frequencies (exact in f32)
    got = _gen(frames, n, sr, wave=fft.WAVE_SINE, freq0=20.0, freq_step=20.0, freq_mod=1000)
    for f in range(frames):
        ref = oracle_mod.osc_sine(sr, float(freqs[f]), n)
        # the phase recurrence is exact on both sides; only sinf differs (CUDA <= 2 ulp, glibc <= 1 ulp)
        assert np.abs(got[f] - ref).max() <= 4e-7, f
    # the piecewise-polynomial waves have no transcendental: bit-identical
    for wave, owave in ((fft.WAVE_FAKE_SINE, 1), (fft.WAVE_SAWTOOTH, 2), (fft.WAVE_SQUARE, 3), (fft.WAVE_TRIANGLE, 4)):
        got = _gen(frames, n, sr, wave=wave, freq0=20.0, freq_step=20.0, freq_mod=1000, duty=0.5)
        for f in range(frames):
            ref = oracle_mod.osc_wave(owave, sr, float(freqs[f]), n, duty=0.5)
            np.testing.assert_array_equal(got[f], ref)
    # configs[0]: 440 Hz at 44.1 kHz; frequency below f32::EPSILON -> silence (src/oscillator.rs:35-37)
    one = _gen(1, 1024, 44100.0, wave=fft.WAVE_SINE, freq0=440.0)[0]
    assert np.abs(one - oracle_mod.osc_sine(44100.0, 440.0, 1024)).max() <= 4e-7
    assert np.all(_gen(2, 64, 48000.0, wave=fft.WAVE_SINE, freq0=0.0) == 0.0)


def test_noise_distributions_and_sharding(built_lib):
    from kopek_b200 import fft
    n, frames = 1024, 512
    u = _gen(frames, n, 48000.0, wave=fft.WAVE_SILENCE, noise=fft.NOISE_RANDOM, seed=7)
    assert u.min() >= -1.0 and u.max() < 1.0                                # random::<f32>() * 2 - 1, src/noise.rs:15-17
    assert abs(u.mean()) < 5e-3 and abs(u.var() - 1.0 / 3.0) < 5e-3
    g = _gen(frames, n, 48000.0, wave=fft.WAVE_SILENCE, noise=fft.NOISE_WHITE, seed=7)
    assert abs(g.mean()) < 5e-3 and abs(g.var() - 1.0) < 1e-2               # StandardNormal, src/noise.rs:19-21
    assert abs((np.abs(g) < 1.0).mean() - 0.6827) < 5e-3 and abs((np.abs(g) < 2.0).mean() - 0.9545) < 3e-3
    # value = osc + gain * noise (examples/beep/generator.rs:62-70); shards reproduce the unsharded signal
    full = _gen(64, n, 48000.0, wave=fft.WAVE_SINE, freq0=100.0, freq_step=50.0, freq_mod=10, noise=fft.NOISE_WHITE,
                noise_gain=0.1, seed=3)
    part = _gen(32, n, 48000.0, wave=fft.WAVE_SINE, freq0=100.0, freq_step=50.0, freq_mod=10, noise=fft.NOISE_WHITE,
                noise_gain=0.1, seed=3, first_frame=32)
    np.testing.assert_array_equal(full[32:], part)
    assert not np.array_equal(full[:32], part)


def test_synth_feeds_the_spectrum_without_host_data(built_lib, oracle_mod):
    """Oscillator -> window -> FFT -> |X| entirely on the device; peak bin where the reference's visual tests expect it."""
    import torch
    from kopek_b200 import StftPlan, OUT_MAG, WINDOW_HANN, fft
    n, frames, sr = 1024, 200, 44100.0
    x = torch.empty((frames, n), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    fft.synth_frames_device(x.data_ptr(), frames, n, sr, wave=fft.WAVE_SINE, freq0=440.0, freq_step=100.0, freq_mod=50,
                            stream=st)
    with StftPlan(n, n, WINDOW_HANN, OUT_MAG) as plan:
        out = torch.empty((frames, plan.bins), device="cuda")
        plan.exec_ptr(x.data_ptr(), frames * n, out.data_ptr(), st)
        torch.cuda.synchronize()
    peaks = out.argmax(dim=1).cpu().numpy()
    want = np.rint((440.0 + 100.0 * (np.arange(frames) % 50)) * n / sr).astype(int)
    assert np.abs(peaks - want).max() <= 1 and peaks[0] == 10          # 440 Hz @ 44.1 kHz -> bin 10


def test_synth_argument_checks(built_lib):
    import torch
    from kopek_b200 import fft, KopekError
    o = torch.empty(1024, device="cuda")
    with pytest.raises(KopekError):
        fft.synth_frames_device(o.data_ptr(), 1, 1024, 48000.0, wave=9)
    with pytest.raises(KopekError):
        fft.synth_frames_device(o.data_ptr(), 1, 1024, 0.0)
    with pytest.raises(KopekError):
        fft.synth_frames_device(o.data_ptr(), 1, 1024, 48000.0, stride=100)
