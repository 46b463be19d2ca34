"""Seeded synthetic inputs for the parity tests and bench.py.

The reference's noise source is an unseeded ThreadRng (src/noise.rs:9-21), so
bit-identical noise is impossible; these generators reproduce its
*distributions* (N(0,1) for white_noise, U[-1,1) for rand_noise) from a seed and
feed the same buffer to the oracle and to the GPU.  Numpy only.
"""
from __future__ import annotations

import numpy as np


def white_noise(count: int, seed: int) -> np.ndarray:
    """N(0,1) f32 -- the distribution of Noise::white_noise (src/noise.rs:19-21)."""
    return np.random.default_rng(seed).standard_normal(count, dtype=np.float32)


def rand_noise(count: int, seed: int) -> np.ndarray:
    """U[-1,1) f32 -- Noise::rand_noise (src/noise.rs:15-17): random::<f32>() * 2.0 - 1.0."""
    return (np.random.default_rng(seed).random(count, dtype=np.float32) * np.float32(2.0) - np.float32(1.0))


def chirp(count: int, sample_rate: float, f0: float, f1: float) -> np.ndarray:
    """Linear chirp f0 -> f1 over the buffer: sin(2 pi (f0 t + (f1-f0) t^2 / (2 T))), f64 then f32."""
    t = np.arange(count, dtype=np.float64) / sample_rate
    dur = count / sample_rate
    return np.sin(2.0 * np.pi * (f0 * t + (f1 - f0) * t * t / (2.0 * dur))).astype(np.float32)


def sine_frames(n_frames: int, n: int, sample_rate: float = 48000.0, noise: float = 0.1, seed: int = 1234,
                first_frame: int = 0) -> np.ndarray:
    """Headline workload M: frame f = sine of 20*(1 + f mod 1000) Hz + noise*N(0,1)."""
    f = (np.arange(first_frame, first_frame + n_frames, dtype=np.float64) % 1000 + 1.0) * 20.0
    t = np.arange(n, dtype=np.float64) / sample_rate
    x = np.sin(2.0 * np.pi * f[:, None] * t[None, :])
    if noise:
        x = x + noise * np.random.default_rng(seed + first_frame).standard_normal((n_frames, n))
    return x.astype(np.float32).reshape(-1)
