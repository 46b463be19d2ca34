"""TEST INFRASTRUCTURE -- ctypes front end of oracle/ref_fft.c.

The CPU restatement of /root/reference/src/fft.rs (PARITY UNPINNED, see the
header of ref_fft.c).  Only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs may import this package; kopek_b200/ must never do so.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libkopek_oracle.so")

WIN_RECT, WIN_HANN, WIN_HAMMING = 0, 1, 2
OUT_MAG, OUT_POWER, OUT_DB = 1, 2, 3


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, no FMA contraction)."""
    src = os.path.join(_HERE, "ref_fft.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        sz, vp, i32, f32, f64 = C.c_size_t, C.c_void_p, C.c_int, C.c_float, C.c_double
        L.oracle_next_pow2.restype = sz
        L.oracle_next_pow2.argtypes = [sz]
        L.oracle_fft_f64.restype = sz
        L.oracle_fft_f64.argtypes = [vp, sz, vp]
        L.oracle_fft_f32.restype = sz
        L.oracle_fft_f32.argtypes = [vp, sz, vp]
        L.oracle_dft_naive_f64.argtypes = [vp, sz, vp]
        L.oracle_show.restype = sz
        L.oracle_show.argtypes = [C.c_char_p, vp, sz, vp, sz]
        L.oracle_osc_sine.argtypes = [f32, f32, vp, vp, sz]
        L.oracle_osc_wave.argtypes = [i32, f32, f32, f32, vp, vp, sz]
        L.oracle_marshal.argtypes = [vp, sz, f64, vp]
        L.oracle_mag_norm_f64.argtypes = [vp, sz, vp]
        L.oracle_mag_sqrt_f32.argtypes = [vp, sz, vp]
        L.oracle_bands_log.argtypes = [vp, vp]
        L.oracle_bands_low.argtypes = [vp, vp]
        L.oracle_peak_band.restype = i32
        L.oracle_peak_band.argtypes = [vp, i32, vp]
        L.oracle_window.argtypes = [i32, sz, vp]
        L.oracle_stft_frames.restype = sz
        L.oracle_stft_frames.argtypes = [sz, sz, sz]
        L.oracle_stft_complex.argtypes = [vp, sz, sz, sz, i32, sz, i32, i32, vp]
        L.oracle_stft_mag.argtypes = [vp, sz, sz, sz, i32, sz, i32, i32, vp]
        L.oracle_max_threads.restype = i32
        L.oracle_pcm16_to_f32.argtypes = [vp, sz, sz, sz, f64, vp]
        L.oracle_pcm16_scale_f32.argtypes = [vp, sz, f32, vp]
        L.oracle_f32_downmix.argtypes = [vp, sz, sz, vp]
        L.oracle_detect_bpm.restype = C.c_int64
        L.oracle_detect_bpm.argtypes = [vp, sz, f32, vp, vp, vp]
        _lib = L
    return _lib


def _p(a: np.ndarray) -> int:
    return a.ctypes.data


def next_pow2(n: int) -> int:
    return int(lib().oracle_next_pow2(n))


def fft(x, dtype=np.complex128) -> np.ndarray:
    """kopek::fft::fft (src/fft.rs:6): complex in, complex out, len next_pow2(len(x))."""
    if dtype == np.complex128:
        a = np.ascontiguousarray(x, dtype=np.complex128)
        out = np.empty(next_pow2(a.size), dtype=np.complex128)
        lib().oracle_fft_f64(_p(a), a.size, _p(out))
    else:
        a = np.ascontiguousarray(x, dtype=np.complex64)
        out = np.empty(next_pow2(a.size), dtype=np.complex64)
        lib().oracle_fft_f32(_p(a), a.size, _p(out))
    return out


def dft_naive(x) -> np.ndarray:
    a = np.ascontiguousarray(x, dtype=np.complex128)
    out = np.empty(a.size, dtype=np.complex128)
    lib().oracle_dft_naive_f64(_p(a), a.size, _p(out))
    return out


def show(label: str, buf) -> str:
    """Text kopek::fft::show (src/fft.rs:43) would print."""
    a = np.ascontiguousarray(buf, dtype=np.complex128)
    need = lib().oracle_show(label.encode(), _p(a), a.size, None, 0)
    dst = C.create_string_buffer(need + 1)
    lib().oracle_show(label.encode(), _p(a), a.size, C.addressof(dst), need + 1)
    return dst.value.decode()


def osc_sine(sample_rate: float, frequency: float, count: int, phase: float = 0.0) -> np.ndarray:
    """Oscillator::sine called `count` times (src/oscillator.rs:47-53)."""
    out = np.empty(count, dtype=np.float32)
    ph = np.array([phase], dtype=np.float32)
    lib().oracle_osc_sine(sample_rate, frequency, _p(ph), _p(out), count)
    return out


def osc_wave(wave: int, sample_rate: float, frequency: float, count: int, duty: float = 0.5) -> np.ndarray:
    out = np.empty(count, dtype=np.float32)
    ph = np.array([0.0], dtype=np.float32)
    lib().oracle_osc_wave(wave, duty, sample_rate, frequency, _p(ph), _p(out), count)
    return out


def marshal(samples, scale_div: float = 1.0) -> np.ndarray:
    s = np.ascontiguousarray(samples, dtype=np.float32)
    out = np.empty(s.size, dtype=np.complex128)
    lib().oracle_marshal(_p(s), s.size, scale_div, _p(out))
    return out


def mag_norm(c) -> np.ndarray:
    a = np.ascontiguousarray(c, dtype=np.complex128)
    out = np.empty(a.size, dtype=np.float64)
    lib().oracle_mag_norm_f64(_p(a), a.size, _p(out))
    return out


def mag_sqrt_f32(c) -> np.ndarray:
    a = np.ascontiguousarray(c, dtype=np.complex128)
    out = np.empty(a.size, dtype=np.float32)
    lib().oracle_mag_sqrt_f32(_p(a), a.size, _p(out))
    return out


def bands_log(y) -> np.ndarray:
    a = np.ascontiguousarray(y, dtype=np.float32)
    assert a.size >= 512
    out = np.empty(8, dtype=np.float32)
    lib().oracle_bands_log(_p(a), _p(out))
    return out


def bands_low(y) -> np.ndarray:
    a = np.ascontiguousarray(y, dtype=np.float32)
    assert a.size >= 44
    out = np.empty(8, dtype=np.float32)
    lib().oracle_bands_low(_p(a), _p(out))
    return out


def peak_band(avg):
    a = np.ascontiguousarray(avg, dtype=np.float32)
    mx = np.zeros(1, dtype=np.float32)
    idx = lib().oracle_peak_band(_p(a), a.size, _p(mx))
    return int(idx), float(mx[0])


def window(kind: int, n: int) -> np.ndarray:
    w = np.empty(n, dtype=np.float32)
    lib().oracle_window(kind, n, _p(w))
    return w


def stft_frames(n_samples: int, n: int, hop: int) -> int:
    return int(lib().oracle_stft_frames(n_samples, n, hop))


def stft_complex(samples, n: int, hop: int, window: int = WIN_RECT, bins: int | None = None,
                 precision: int = 64, threads: int = 0) -> np.ndarray:
    s = np.ascontiguousarray(samples, dtype=np.float32)
    bins = n // 2 + 1 if bins is None else bins
    frames = stft_frames(s.size, n, hop)
    out = np.empty((frames, bins), dtype=np.complex128)
    lib().oracle_stft_complex(_p(s), s.size, n, hop, window, bins, precision, threads, _p(out))
    return out


def stft_mag(samples, n: int, hop: int, window: int = WIN_RECT, bins: int | None = None,
             output: int = OUT_MAG, threads: int = 0) -> np.ndarray:
    s = np.ascontiguousarray(samples, dtype=np.float32)
    bins = n // 2 + 1 if bins is None else bins
    frames = stft_frames(s.size, n, hop)
    out = np.empty((frames, bins), dtype=np.float32)
    lib().oracle_stft_mag(_p(s), s.size, n, hop, window, bins, output, threads, _p(out))
    return out


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def pcm16_to_f32(pcm, channel: int = 0, divisor: float = 32767.0) -> np.ndarray:
    """examples/graph/player.rs:56-59: frame[channel] as f64 / i16::MAX, narrowed once to f32."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    if pcm.ndim == 1:
        pcm = pcm[:, None]
    out = np.empty(pcm.shape[0], dtype=np.float32)
    lib().oracle_pcm16_to_f32(_p(pcm), pcm.shape[0], pcm.shape[1], channel, divisor, _p(out))
    return out


def pcm16_scale_f32(pcm, factor: float = 0.00002) -> np.ndarray:
    """examples/graph/player.rs:81-102 (load_track_at_path): volume_factor * (value as f32), channels interleaved."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1)
    out = np.empty(pcm.size, dtype=np.float32)
    lib().oracle_pcm16_scale_f32(_p(pcm), pcm.size, factor, _p(out))
    return out


def f32_downmix(frames, channels: int) -> np.ndarray:
    """examples/pong/audio.rs:57-70: (0.0 + f[0] + f[1]) / channels per interleaved frame."""
    f = np.ascontiguousarray(frames, dtype=np.float32).reshape(-1, channels)
    out = np.empty(f.shape[0], dtype=np.float32)
    lib().oracle_f32_downmix(_p(f), f.shape[0], channels, _p(out))
    return out


def detect_bpm(frames, sample_rate: float = 44100.0):
    """src/decoder.rs:39-71.  frames: [n][2] int16.  Returns (bpm, beats, block_e[duration*43], avg_e[duration]);
    bpm is None where the Rust original panics (duration 0, or fewer than duration*44100 frames)."""
    frames = np.ascontiguousarray(frames, dtype=np.int16).reshape(-1, 2)
    duration = int(np.float32(frames.shape[0]) / np.float32(sample_rate))
    block_e = np.zeros(max(duration, 0) * 43, dtype=np.float32)
    avg_e = np.zeros(max(duration, 0), dtype=np.float32)
    beats = C.c_uint32(0)
    bpm = lib().oracle_detect_bpm(_p(frames), frames.shape[0], sample_rate, C.byref(beats), _p(block_e), _p(avg_e))
    return (None if bpm < 0 else int(bpm)), int(beats.value), block_e, avg_e
