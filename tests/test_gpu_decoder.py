"""GPU: the decoder-side callers through the C ABI against the oracle -- bit-exact.

kopek_pcm16_to_f32 = examples/graph/player.rs:56-59 (frame[0] as f64 / i16::MAX); kopek_detect_bpm = src/decoder.rs:39-71."""
import numpy as np
import pytest

from test_oracle import _click_track
from util import frame_rel_err, tol

pytestmark = pytest.mark.gpu


def test_pcm16_to_f32_is_bit_exact(built_lib, oracle_mod):
    import torch
    from kopek_b200 import fft, KopekError
    rng = np.random.default_rng(3)
    for channels, n, offset in ((2, 100003, 0), (2, 4099, 1), (1, 5000, 0), (3, 777, 0), (2, 3, 0)):
        pcm = rng.integers(-32768, 32768, size=(n + offset, channels), dtype=np.int16)
        pcm[offset:offset + 3, 0] = [32767, -32768, 0]
        d = torch.from_numpy(pcm).cuda()
        for ch in range(channels):
            out = torch.full((n + 8,), float("nan"), device="cuda")
            fft.pcm16_to_f32_device(d.data_ptr() + offset * channels * 2, n, out.data_ptr(), channels, ch)
            torch.cuda.synchronize()
            got = out.cpu().numpy()
            np.testing.assert_array_equal(got[:n], oracle_mod.pcm16_to_f32(pcm[offset:], ch))
            assert np.isnan(got[n:]).all()
    with pytest.raises(KopekError):
        fft.pcm16_to_f32_device(d.data_ptr(), 3, out.data_ptr(), 2, 2)          # channel out of range
    with pytest.raises(KopekError):
        fft.pcm16_to_f32_device(d.data_ptr(), 3, out.data_ptr(), 2, 0, divisor=0.0)


@pytest.mark.parametrize("seconds,bpm,seed", [(3.2, 120, 0), (7.9, 90, 1), (1.0, 60, 2), (12.5, 174, 3), (61.5, 128, 4),
                                              (300.0, 100, 5)])
def test_detect_bpm_is_bit_exact(built_lib, oracle_mod, seconds, bpm, seed):
    import torch
    from kopek_b200 import fft
    x = _click_track(seconds, bpm, seed)
    ref_bpm, ref_beats, ref_block, ref_avg = oracle_mod.detect_bpm(x)
    d = torch.from_numpy(x).cuda()
    be = torch.full((ref_block.size + 4,), float("nan"), device="cuda")
    av = torch.full((ref_avg.size + 4,), float("nan"), device="cuda")
    got_bpm, got_beats = fft.detect_bpm_device(d.data_ptr(), x.shape[0], 44100.0, be.data_ptr(), av.data_ptr())
    assert (got_bpm, got_beats) == (ref_bpm, ref_beats)
    np.testing.assert_array_equal(be.cpu().numpy()[:ref_block.size], ref_block)
    np.testing.assert_array_equal(av.cpu().numpy()[:ref_avg.size], ref_avg)
    assert np.isnan(be.cpu().numpy()[ref_block.size:]).all() and np.isnan(av.cpu().numpy()[ref_avg.size:]).all()
    # without the optional outputs
    assert fft.detect_bpm_device(d.data_ptr(), x.shape[0]) == (ref_bpm, ref_beats)
    if seconds > 60:
        assert got_bpm == 0 and got_beats > 0        # 60 / duration == 0 in u32 arithmetic (decoder.rs:70)


def test_detect_bpm_rejects_what_the_reference_panics_on(built_lib):
    import torch
    from kopek_b200 import fft, KopekError
    d = torch.zeros((50000, 2), dtype=torch.int16, device="cuda")
    with pytest.raises(KopekError):
        fft.detect_bpm_device(d.data_ptr(), 30000)                  # duration 0 -> 60 / 0
    with pytest.raises(KopekError):
        fft.detect_bpm_device(d.data_ptr(), 50000, 22050.0)         # 2 "seconds" x 44100 frames > 50000
    assert fft.detect_bpm_device(d.data_ptr(), 50000) == (0, 0)     # silence: 0 > 0 is false


def test_decoded_frames_to_spectrum(built_lib, oracle_mod):
    """The graph player's path end to end (examples/graph/player.rs:56-61, utils.rs:47-63): i16 frames -> channel 0 /
    32767 -> fft -> magnitude, on device, against the f64 oracle."""
    import torch
    from kopek_b200 import fft, StftPlan, OUT_MAG, WINDOW_RECT
    n, frames = 1024, 37
    x = _click_track(1.0, 140, 7)[: n * frames]
    d = torch.from_numpy(x).cuda()
    samples = torch.empty(n * frames, device="cuda")
    fft.pcm16_to_f32_device(d.data_ptr(), n * frames, samples.data_ptr(), 2, 0)
    out = torch.empty((frames, n // 2 + 1), device="cuda")
    with StftPlan(n, n, WINDOW_RECT, OUT_MAG) as plan:
        plan.exec_ptr(samples.data_ptr(), n * frames, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = np.stack([oracle_mod.mag_norm(oracle_mod.fft(oracle_mod.marshal(x[f * n:(f + 1) * n, 0].astype(np.float32), 32767.0)))
                    [: n // 2 + 1] for f in range(frames)])
    assert frame_rel_err(out.cpu().numpy().astype(np.float64), ref) <= tol(n)


def test_detect_bpm_host_entry_point(built_lib, oracle_mod):
    """kopek_detect_bpm_host: the drop-in shape of detect_bpm(frames: Vec<[i16; 2]>, sample_rate) -> u32."""
    from kopek_b200 import fft, KopekError
    for seconds, bpm, seed in ((2.82, 83, 11), (9.3, 140, 12), (400.0, 120, 13)):
        x = _click_track(seconds, bpm, seed)
        ref_bpm, ref_beats, _, _ = oracle_mod.detect_bpm(x)
        assert fft.detect_bpm(x) == (ref_bpm, ref_beats)
    with pytest.raises(KopekError):
        fft.detect_bpm(_click_track(0.3, 120, 14)[:10000])


@pytest.mark.parametrize("n,hop,output,window", [(1024, 1024, 1, 0), (1024, 256, 3, 1), (1024, 1024, 0, 2), (1024, 512, 2, 1),
                                                 (1024, 1024, 4, 0), (1024, 1024, 5, 1), (2048, 512, 1, 1),
                                                 (1024, 333, 1, 1), (256, 256, 1, 0), (4096, 4096, 3, 2)])
def test_pcm16_frames_straight_into_the_plan(built_lib, oracle_mod, n, hop, output, window):
    """kopek_stft_exec_pcm16: fused marshalling (n = 1024) or conversion + plan (elsewhere) -- bit-identical to
    kopek_pcm16_to_f32 followed by kopek_stft_exec, and within tolerance of the f64 oracle on the marshalled samples."""
    import torch
    from kopek_b200 import fft, StftPlan
    frames = 301
    n_frames = (frames - 1) * hop + n + 5
    x = _click_track(n_frames / 44100.0 + 0.1, 150, n + hop)[:n_frames]
    x[:4, 0] = [32767, -32768, 0, -1]
    x[:4, 1] = [-32768, 32767, 1, 0]
    d = torch.from_numpy(x).cuda()
    st = torch.cuda.current_stream().cuda_stream
    for channel in (0, 1):
        with StftPlan(n, hop, window, output) as plan:
            dt = torch.complex64 if output == 0 else torch.float32
            a = torch.full((frames, plan.pitch), float("nan"), device="cuda", dtype=dt)
            b = torch.full((frames, plan.pitch), float("nan"), device="cuda", dtype=dt)
            assert plan.exec_pcm16_ptr(d.data_ptr(), n_frames, a.data_ptr(), channel, st) == frames
            assert plan.last_launches == (1 if n == 1024 and hop % 4 == 0 else 2)
            samples = torch.empty(n_frames, device="cuda")
            fft.pcm16_to_f32_device(d.data_ptr(), n_frames, samples.data_ptr(), 2, channel)
            assert plan.exec_ptr(samples.data_ptr(), n_frames, b.data_ptr(), st) == frames
            torch.cuda.synchronize()
            ra, rb = (torch.view_as_real(a), torch.view_as_real(b)) if output == 0 else (a, b)
            assert torch.equal(ra, rb)
            if output == 1:
                ref = oracle_mod.stft_mag(oracle_mod.pcm16_to_f32(x, channel), n, hop, window)
                assert (np.abs(a.cpu().numpy() - ref) / ref.max(axis=1, keepdims=True)).max() <= tol(n)


def test_pcm16_scale_f32_is_bit_exact(built_lib, oracle_mod):
    """load_track_at_path (examples/graph/player.rs:81-102): volume_factor * (value as f32) over the interleaved frames;
    every i16 value, aligned and unaligned buffers, lengths that leave a tail, the 128-bit path and the scalar one."""
    import torch
    from kopek_b200 import fft
    allv = np.arange(-32768, 32768, dtype=np.int16)                       # every possible value once
    rng = np.random.default_rng(9)
    for pcm, offset in ((allv, 0), (rng.integers(-32768, 32768, size=200_003, dtype=np.int16), 0),
                        (rng.integers(-32768, 32768, size=4099, dtype=np.int16), 1), (allv[:5], 0)):
        d = torch.from_numpy(pcm).cuda()
        n = pcm.size - offset
        for factor in (0.00002, 1.0 / 32767.0, -3.5):
            out = torch.full((n + 8,), float("nan"), device="cuda")
            fft.pcm16_scale_f32_device(d.data_ptr() + 2 * offset, n, factor, out.data_ptr())
            torch.cuda.synchronize()
            got = out.cpu().numpy()
            want = oracle_mod.pcm16_scale_f32(pcm[offset:], factor)
            assert np.array_equal(got[:n].view(np.uint32), want.view(np.uint32)), (pcm.size, offset, factor)
            assert np.isnan(got[n:]).all()
    # the numpy statement of the same expression (f32 product) agrees with the oracle
    want = np.float32(0.00002) * allv.astype(np.float32)
    assert np.array_equal(oracle_mod.pcm16_scale_f32(allv, 0.00002).view(np.uint32), want.view(np.uint32))


def test_f32_downmix_is_bit_exact(built_lib, oracle_mod):
    """The pong capture callback (examples/pong/audio.rs:57-70): first two channels summed from 0.0, divided by the channel
    count -- mono, stereo (aligned and not), four channels; signed zeros, infinities and NaN go the reference's way."""
    import torch
    from kopek_b200 import fft
    rng = np.random.default_rng(4)
    for channels, n, offset in ((2, 100_001, 0), (2, 5000, 1), (1, 4097, 0), (4, 3001, 0), (3, 7, 0)):
        f = rng.standard_normal((n + offset, channels)).astype(np.float32)
        f[offset:offset + 4, 0] = [-0.0, np.inf, np.nan, 1e-45]
        f[offset:offset + 4, -1] = [-0.0, -np.inf, 1.0, -1e-45]
        d = torch.from_numpy(f).cuda()
        out = torch.full((n + 8,), 7.0, device="cuda")
        fft.f32_downmix_device(d.data_ptr() + 4 * channels * offset, n, channels, out.data_ptr())
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        want = oracle_mod.f32_downmix(f[offset:], channels)
        nan = np.isnan(want)
        assert np.array_equal(np.isnan(got[:n]), nan)                        # NaN where the reference gives NaN (payloads differ by ISA)
        assert np.array_equal(got[:n][~nan].view(np.uint32), want[~nan].view(np.uint32)), (channels, n, offset)
        assert (got[n:] == 7.0).all()
