"""kopek_b200 -- B200-native FFT frequency-analysis path of zehreken/kopek.

Layout:
  csrc/        hand-written sm_100a CUDA kernels + the C ABI (include/kopek_cuda.h)
  lib/         libkopek_cuda.so, built in-tree by `python -m kopek_b200.build`
  fft.py       host mirror of the reference's `kopek::fft` module (fft, show) + StftPlan
  sharded.py   frame-range sharding over one process per GPU (+ opt-in gather)
  signals.py   seeded synthetic inputs (sine / chirp / noise) for tests and bench
"""
from . import _lib  # noqa: F401
from . import fft  # noqa: F401  -- kopek_b200.fft.fft / kopek_b200.fft.show mirror kopek::fft::{fft, show}
from .fft import (BAND_ROW, BINS_FULL, BINS_HALF, OUT_BANDS_LOG, OUT_BANDS_LOW, OUT_COMPLEX, OUT_DB, OUT_MAG, OUT_POWER, WINDOW_CUSTOM, WINDOW_HAMMING,
                  WINDOW_HANN, WINDOW_RECT, KopekError, LiveSpectrum, StftPlan)

__version__ = "0.1.0"
