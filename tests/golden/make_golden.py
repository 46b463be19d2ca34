#!/usr/bin/env python3
"""Regenerate tests/golden/*.npz.

The reference (Rust) cannot be executed in this image and holds no golden
vector for src/fft.rs, so these fixtures pin the oracle and the GPU path
against an INDEPENDENT implementation (numpy.fft, pocketfft, f64) on inputs
built the way the reference's callers build them:
  * sine440:  Oscillator::sine at 44.1 kHz / 440 Hz (src/oscillator.rs:47-53, via the oracle's
              restatement of the f32 phase accumulator), 1024 samples -- BASELINE.json configs[0]
  * noise256: seeded N(0,1) frames (distribution of src/noise.rs:19-21) -- configs[2]
  * chirp:    48 kHz linear chirp 20 Hz -> 20 kHz, Hann 2048 / hop 512, selected frames -- configs[1]
  * ragged:   1000 complex points (zero-padded to 1024 by fft, src/fft.rs:31-36)
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402
from kopek_b200 import signals  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    # configs[0]
    s = oracle.osc_sine(44100.0, 440.0, 1024)
    X = np.fft.fft(s.astype(np.float64))
    np.savez_compressed(os.path.join(HERE, "sine440_1024.npz"), samples=s, spectrum=X,
                        peak_bin=np.int64(np.abs(X[:513]).argmax()), peak_mag=np.abs(X).max())
    # configs[2] (4 frames)
    x = signals.white_noise(4 * 256, seed=42)
    np.savez_compressed(os.path.join(HERE, "noise256.npz"), samples=x,
                        spectrum=np.fft.rfft(x.astype(np.float64).reshape(4, 256), axis=1))
    # configs[1] (frames 0, 1000, 5621 of 5622)
    n, hop = 2048, 512
    c = signals.chirp(60 * 48000, 48000.0, 20.0, 20000.0)
    w = (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)).astype(np.float32)
    idx = np.array([0, 1000, 5621])
    fr = np.stack([c[i * hop:i * hop + n] for i in idx])
    xw = (fr * w[None, :]).astype(np.float32)          # f32 product, as the fused path defines it
    np.savez_compressed(os.path.join(HERE, "chirp_hann2048.npz"), frames=fr, frame_index=idx,
                        spectrum=np.fft.rfft(xw.astype(np.float64), axis=1))
    # ragged complex input
    rng = np.random.default_rng(5)
    z = rng.standard_normal(1000) + 1j * rng.standard_normal(1000)
    np.savez_compressed(os.path.join(HERE, "ragged1000.npz"), input=z, spectrum=np.fft.fft(z, 1024))
    # inputs of the Rust pinning kit (rust/golden): interleaved (re, im) f64, little endian -- what the reference's
    # callers hand to fft(): f32 samples widened to f64, im = 0 (examples/beep/view.rs:140-144)
    for name, arr in (("sine440", oracle.marshal(s)),                                  # configs[0], 1024 points
                      ("ragged1000", z),                                               # zero-padded to 1024
                      ("noise256", oracle.marshal(x[:256])),                           # one configs[2] frame
                      ("len0", np.zeros(0, np.complex128)), ("len1", np.array([2.5 - 1.5j])), ("len3", z[:3]),
                      ("pcm4096", oracle.marshal(signals.chirp(4096, 48000.0, 20.0, 20000.0), 32767.0))):   # the /32767 quirk
        np.ascontiguousarray(arr, dtype="<c16").tofile(os.path.join(HERE, f"rust_in_{name}.bin"))
    # show() text
    txt = "label\n1.0000+2.0000i, -0.5000-0.2500i, 0.0000+0.0000i, -0.0000-3.1416i\n"
    with open(os.path.join(HERE, "show.txt"), "w") as f:
        f.write(txt)


if __name__ == "__main__":
    main()
