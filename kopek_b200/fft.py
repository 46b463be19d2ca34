"""Host-side mirror of the reference's `kopek::fft` module over the C ABI.

Same names, argument meaning and error behaviour as /root/reference/src/fft.rs:
`fft(input)` takes complex samples and returns the forward DFT zero-padded to
the next power of two (src/fft.rs:6-41); `show(label, buf)` prints the bins in
the reference's format (src/fft.rs:43-52).  The reference's `fft` is infallible,
so any CUDA failure surfaces as an exception (the Rust shim panics) -- there is
no CPU path to fall back to.

`StftPlan` is the batched throughput path the three callers of `fft`
(examples/beep/view.rs:140-158, examples/graph/player.rs:56-61,
examples/pong/main.rs:192-196) map onto when many frames are analysed at once.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import (BAND_ROW, BINS_FULL, BINS_HALF, OUT_BANDS_LOG, OUT_BANDS_LOW, OUT_COMPLEX, OUT_DB, OUT_MAG, OUT_POWER, WINDOW_CUSTOM, WINDOW_HAMMING,
                   WINDOW_HANN, WINDOW_RECT, KopekError, check)

__all__ = ["fft", "fft_batched", "fft_large_device", "synth_frames_device", "LiveSpectrum", "show", "show_format", "StftPlan", "next_pow2", "KopekError",
           "WINDOW_RECT", "WINDOW_HANN", "WINDOW_HAMMING", "WINDOW_CUSTOM",
           "OUT_COMPLEX", "OUT_MAG", "OUT_POWER", "OUT_DB", "OUT_BANDS_LOG", "OUT_BANDS_LOW", "BAND_ROW",
           "BINS_HALF", "BINS_FULL"]


def next_pow2(n: int) -> int:
    return int(_lib.load().kopek_next_pow2(n))


def fft(input) -> np.ndarray:
    """kopek::fft::fft -- complex128 in, complex128 out of length next_pow2(len(input)) (>= 1)."""
    a = np.ascontiguousarray(input, dtype=np.complex128).reshape(-1)
    out = np.empty(next_pow2(a.size), dtype=np.complex128)
    check(_lib.load().kopek_fft_c2c_f64(a.ctypes.data if a.size else None, a.size, out.ctypes.data, out.size))
    return out


def fft_batched(frames) -> np.ndarray:
    """`fft` over each row of a [batch, n_in] complex array in one call."""
    a = np.ascontiguousarray(frames, dtype=np.complex128)
    if a.ndim != 2:
        raise ValueError("frames must be 2-D [batch, n]")
    out = np.empty((a.shape[0], next_pow2(a.shape[1])), dtype=np.complex128)
    check(_lib.load().kopek_fft_c2c_f64_batched(a.ctypes.data if a.size else None, a.shape[1], a.shape[0],
                                                out.ctypes.data))
    return out


def fft_large_device(d_in: int, d_out: int, n: int, device: int = 0, stream: int = 0) -> None:
    """Four-step complex f32 FFT of one long device-resident buffer (kopek_fft_large_c2c_f32):
    n interleaved (re,im) pairs at d_in -> n at d_out, natural order, asynchronous on `stream`."""
    check(_lib.load().kopek_fft_large_c2c_f32(d_in, d_out, n, device, stream or None))


WAVE_SINE, WAVE_FAKE_SINE, WAVE_SAWTOOTH, WAVE_SQUARE, WAVE_TRIANGLE, WAVE_SILENCE = 0, 1, 2, 3, 4, 5
NOISE_NONE, NOISE_RANDOM, NOISE_WHITE = 0, 1, 2


def synth_frames_device(d_out: int, n_frames: int, n: int, sample_rate: float, wave: int = WAVE_SINE, freq0: float = 440.0,
                        freq_step: float = 0.0, freq_mod: int = 0, noise: int = NOISE_NONE, noise_gain: float = 1.0,
                        seed: int = 0, first_frame: int = 0, duty: float = 0.5, stride: int | None = None,
                        device: int = 0, stream: int = 0) -> None:
    """kopek_synth_frames: Oscillator (+ Noise) frames generated on the device (src/oscillator.rs, src/noise.rs)."""
    check(_lib.load().kopek_synth_frames(d_out, n_frames, n, n if stride is None else stride, sample_rate, wave, duty,
                                         freq0, freq_step, freq_mod, noise, noise_gain, seed, first_frame, device,
                                         stream or None))


def ifft(spectrum) -> np.ndarray:
    """Inverse through the forward drop-in: conj(fft(conj(X))) / n, n a power of two (kopek_ifft_c2c_f64)."""
    a = np.ascontiguousarray(spectrum, dtype=np.complex128).reshape(-1)
    out = np.empty(a.size, dtype=np.complex128)
    check(_lib.load().kopek_ifft_c2c_f64(a.ctypes.data, a.size, out.ctypes.data))
    return out


def pcm16_to_f32_device(d_pcm: int, n_frames: int, d_out: int, channels: int = 2, channel: int = 0,
                        divisor: float = 32767.0, device: int = 0, stream: int = 0) -> None:
    """kopek_pcm16_to_f32: frame[channel] as f64 / i16::MAX (examples/graph/player.rs:56-59), device pointers."""
    check(_lib.load().kopek_pcm16_to_f32(d_pcm, n_frames, channels, channel, divisor, d_out, device, stream or None))


def pcm16_scale_f32_device(d_pcm: int, n_values: int, factor: float, d_out: int, device: int = 0, stream: int = 0) -> None:
    """kopek_pcm16_scale_f32: volume_factor * (value as f32) over interleaved i16 values (examples/graph/player.rs:81-102)."""
    check(_lib.load().kopek_pcm16_scale_f32(d_pcm, n_values, factor, d_out, device, stream or None))


def f32_downmix_device(d_frames: int, n_frames: int, channels: int, d_out: int, device: int = 0, stream: int = 0) -> None:
    """kopek_f32_downmix: (f[0] + f[1]) / channels per interleaved frame (examples/pong/audio.rs:57-70)."""
    check(_lib.load().kopek_f32_downmix(d_frames, n_frames, channels, d_out, device, stream or None))


def detect_bpm_device(d_frames: int, n_frames: int, sample_rate: float = 44100.0, d_block_e: int = 0, d_avg_e: int = 0,
                      device: int = 0, stream: int = 0):
    """kopek_detect_bpm: kopek::decoder::detect_bpm (src/decoder.rs:39-71) on device-resident [n][2] i16 frames.
    Returns (bpm, beats)."""
    bpm, beats = C.c_uint32(0), C.c_uint32(0)
    check(_lib.load().kopek_detect_bpm(d_frames, n_frames, sample_rate, C.byref(bpm), C.byref(beats), d_block_e or None,
                                       d_avg_e or None, device, stream or None))
    return int(bpm.value), int(beats.value)


def detect_bpm(frames, sample_rate: float = 44100.0, device: int = 0):
    """kopek::decoder::detect_bpm(frames: Vec<[i16; 2]>, sample_rate) -- host frames in, (bpm, beats) out."""
    f = np.ascontiguousarray(frames, dtype=np.int16).reshape(-1, 2)
    bpm, beats = C.c_uint32(0), C.c_uint32(0)
    check(_lib.load().kopek_detect_bpm_host(f.ctypes.data, f.shape[0], sample_rate, C.byref(bpm), C.byref(beats), device))
    return int(bpm.value), int(beats.value)


def show_format(label: str, buf) -> str:
    a = np.ascontiguousarray(buf, dtype=np.complex128).reshape(-1)
    lib = _lib.load()
    need = lib.kopek_show_format(label.encode(), a.ctypes.data, a.size, None, 0)
    dst = C.create_string_buffer(need + 1)
    lib.kopek_show_format(label.encode(), a.ctypes.data, a.size, C.addressof(dst), need + 1)
    return dst.value.decode()


def show(label: str, buf) -> None:
    """kopek::fft::show -- prints the label line and the "{:.4}{:+.4}i" bins."""
    a = np.ascontiguousarray(buf, dtype=np.complex128).reshape(-1)
    check(_lib.load().kopek_show(label.encode(), a.ctypes.data, a.size))


class StftPlan:
    """Fused window -> real FFT -> |X|/dB plan (kopek_plan_* / kopek_stft_exec*)."""

    def __init__(self, n: int, hop: int | None = None, window: int = WINDOW_RECT, output: int = OUT_MAG,
                 bins: int = BINS_HALF, out_pitch: int = 0, device: int = 0, custom_window=None,
                 band_scale: float | None = None, bulk_store: bool = False):
        self._h = C.c_void_p()
        self._lib = _lib.load()
        hop = n if hop is None else hop
        check(self._lib.kopek_plan_create(C.byref(self._h), n, hop, window, output, bins, out_pitch, device))
        self.n, self.hop, self.window, self.output, self.device = n, hop, window, output, device
        self.bins = int(self._lib.kopek_plan_bins(self._h))
        self.pitch = int(self._lib.kopek_plan_out_pitch(self._h))
        self.elem_bytes = int(self._lib.kopek_plan_out_elem_bytes(self._h))
        if bulk_store:
            check(self._lib.kopek_plan_set_bulk_store(self._h, 1))
        if band_scale is not None:
            check(self._lib.kopek_plan_set_band_scale(self._h, band_scale))
        if custom_window is not None:
            w = np.ascontiguousarray(custom_window, dtype=np.float32)
            if w.size != n:
                raise ValueError("custom window must have n coefficients")
            check(self._lib.kopek_plan_set_window(self._h, w.ctypes.data))

    def frames(self, n_samples: int) -> int:
        return int(self._lib.kopek_plan_frames(self._h, n_samples))

    @property
    def last_launches(self) -> int:
        return int(self._lib.kopek_plan_last_launches(self._h))

    @property
    def out_dtype(self):
        return np.complex64 if self.output == OUT_COMPLEX else np.float32

    def exec_ptr(self, d_samples: int, n_samples: int, d_out: int, stream: int = 0) -> int:
        """Device pointers in, asynchronous on `stream` (a cudaStream_t as int)."""
        nf = C.c_uint64(0)
        check(self._lib.kopek_stft_exec(self._h, d_samples, n_samples, d_out, C.byref(nf), stream or None))
        return nf.value

    def exec_pcm16_ptr(self, d_frames: int, n_frames: int, d_out: int, channel: int = 0, stream: int = 0) -> int:
        """Interleaved stereo i16 frames (device) in: frame[channel] / i16::MAX feeds the plan (fused for n = 1024)."""
        nf = C.c_uint64(0)
        check(self._lib.kopek_stft_exec_pcm16(self._h, d_frames, n_frames, channel, d_out, C.byref(nf), stream or None))
        return nf.value

    def exec_host_ptr(self, h_samples: int, n_samples: int, h_out: int) -> int:
        nf = C.c_uint64(0)
        check(self._lib.kopek_stft_exec_host(self._h, h_samples, n_samples, h_out, C.byref(nf)))
        return nf.value

    def exec_host(self, samples) -> np.ndarray:
        """Host numpy samples in, [frames, bins] spectra out (copies inside the call)."""
        s = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1)
        frames = self.frames(s.size)
        out = np.empty((frames, self.pitch), dtype=self.out_dtype)
        if frames:
            self.exec_host_ptr(s.ctypes.data, s.size, out.ctypes.data)
        return out[:, :self.bins]

    def close(self) -> None:
        if self._h:
            self._lib.kopek_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class LiveSpectrum:
    """Sliding-window analyser of the live examples (kopek_stream_*): the VecDeque of the most recent n samples
    of examples/beep/view.rs:44,57-60 kept in mapped pinned memory the kernel reads directly; `push(new_samples)` returns the spectrum of the
    window that ends at the newest sample (history starts as zeros)."""

    def __init__(self, n: int = 1024, window: int = WINDOW_RECT, output: int = OUT_MAG, input_scale: float = 1.0,
                 device: int = 0):
        self._h = C.c_void_p()
        self._lib = _lib.load()
        check(self._lib.kopek_stream_create(C.byref(self._h), n, window, output, input_scale, device))
        self.n, self.output = n, output
        self.bins = int(self._lib.kopek_stream_bins(self._h))

    def push(self, samples) -> np.ndarray:
        s = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1)
        out = np.empty(self.bins, dtype=np.complex64 if self.output == OUT_COMPLEX else np.float32)
        check(self._lib.kopek_stream_push(self._h, s.ctypes.data if s.size else None, s.size, out.ctypes.data))
        return out

    def close(self) -> None:
        if self._h:
            self._lib.kopek_stream_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
