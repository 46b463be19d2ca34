#!/usr/bin/env python3
"""Numpy model of the 8-lanes-per-frame 256-point kernel (fft_warp256.cuh): M = 128 = 8 x 16, four frames per warp.
Checks the spectrum against numpy.fft.rfft and counts shared-memory wavefronts (design tool)."""
import numpy as np

from k1w_model import Smem, cw

M, N = 128, 256


def run_warp(x4frames, win=None):
    sm = Smem(288 * 16)
    lanes = range(32)
    w4 = (0.5 * win).reshape(64, 4) if win is not None else None
    dA = {}
    for l in lanes:
        g, a = l >> 3, l & 7
        z4 = x4frames[g].reshape(64, 4)
        d = np.zeros((2, 8), complex)
        for n1 in range(8):
            v = z4[8 * n1 + a] * (w4[8 * n1 + a] if w4 is not None else 1.0)
            d[0, n1] = v[0] + 1j * v[1]
            d[1, n1] = v[2] + 1j * v[3]
        A = np.stack([np.fft.fft(d[0]), np.fft.fft(d[1])])
        for k1 in range(1, 8):
            A[0, k1] *= cw(2 * a * k1, 128)
            A[1, k1] *= cw((2 * a + 1) * k1, 128)
        dA[l] = A
    for k1 in range(8):
        sm.access(f"A.sts k1={k1}", {l: 16 * (72 * (l >> 3) + 9 * k1 + (l & 7)) for l in lanes}, 16, True,
                  {l: [dA[l][0, k1].real, dA[l][0, k1].imag, dA[l][1, k1].real, dA[l][1, k1].imag] for l in lanes})
    Z = {}
    e = {l: np.zeros(16, complex) for l in lanes}
    for s in range(8):
        got = sm.access(f"B.lds s={s}", {l: 16 * (72 * (l >> 3) + 9 * (l & 7) + s) for l in lanes}, 16, False)
        for l in lanes:
            e[l][2 * s] = got[l][0] + 1j * got[l][1]
            e[l][2 * s + 1] = got[l][2] + 1j * got[l][3]
    for l in lanes:
        Z[l] = np.fft.fft(e[l])              # Z[a + 8 k0]
    for j in range(4):
        sm.access(f"Z.sts j={j}", {l: 16 * (72 * (l >> 3) + (l & 7) + 8 * j) for l in lanes}, 16, True,
                  {l: [Z[l][8 + 2 * j].real, Z[l][8 + 2 * j].imag, Z[l][9 + 2 * j].real, Z[l][9 + 2 * j].imag] for l in lanes})
    X = np.full((4, M + 1), np.nan, complex)
    scale = 1.0 if win is not None else 0.5
    pw = lambda k: complex(-np.sin(2 * np.pi * k / N), -np.cos(2 * np.pi * k / N))
    for i in range(4):
        got = sm.access(f"P.lds i={i}", {l: 16 * (72 * (l >> 3) + ((8 - (l & 7)) & 7) + 8 * (3 - i)) for l in lanes}, 16, False)
        addr2 = {l: 16 * (72 * (l >> 3) + 0 + 8 * (4 - i)) for l in lanes if (l & 7) == 0 and i >= 1}
        got2 = sm.access(f"P2.lds i={i}", addr2, 16, False) if addr2 else {}
        for l in lanes:
            g, a = l >> 3, l & 7
            lo = got[l][0] + 1j * got[l][1]
            hi = got[l][2] + 1j * got[l][3]
            for par in range(2):
                k0 = 2 * i + par
                k = a + 8 * k0
                A_ = Z[l][k0]
                if a >= 1:
                    B_ = hi if par == 0 else lo
                elif par == 1:
                    B_ = hi
                elif i == 0:
                    B_ = A_
                else:
                    B_ = got2[l][0] + 1j * got2[l][1]
                S = A_ + np.conj(B_)
                D = A_ - np.conj(B_)
                Q = D * pw(k)
                X[g, k] = (S + Q) * scale
                X[g, M - k] = np.conj(S - Q) * scale
    for g in range(4):
        X[g, 64] = np.conj(Z[8 * g][8]) * 2 * scale
    return X, sm.log


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    x = rng.standard_normal((4, N))
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / N)
    for win in (None, w):
        X, log = run_warp(x, win)
        ref = np.fft.rfft(x * (win if win is not None else 1.0), axis=1)
        print("max rel err", np.abs(X - ref).max() / np.abs(ref).max())
    by = {}
    for name, wf, width, st in log:
        k = name.split()[0]
        by.setdefault(k, [0, 0])
        by[k][0] += wf
        by[k][1] += 1
    for k, (wf, n) in by.items():
        print(f"{k:8s} instr {n:3d} wavefronts {wf:4d} (per 4 frames)")
