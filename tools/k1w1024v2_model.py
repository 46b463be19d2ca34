#!/usr/bin/env python3
"""Numpy model of the 2.5-pass 1024-point kernel (fft_warp1024b.cuh): M = 512 = 16 x 32 on 32 lanes.
Pass A: radix-16 over n1 (lane = n0).  Pass B: lane pair (k1, par) -> radix-16 over the n0 of its parity, then the
last radix-2 across the pair by shuffles.  Checks against numpy and counts shared wavefronts (design tool)."""
import numpy as np

from k1w_model import Smem, cw

M, N = 512, 1024


def run_frame(x, win=None):
    sm = Smem(544 * 8)
    lanes = range(32)
    z2 = x.reshape(512, 2)
    w2 = (0.5 * win).reshape(512, 2) if win is not None else None
    # ---- pass A
    U = {}
    for l in lanes:
        v = np.zeros(16, complex)
        for n1 in range(16):
            s = z2[32 * n1 + l] * (w2[32 * n1 + l] if w2 is not None else 1.0)
            v[n1] = s[0] + 1j * s[1]
        u = np.fft.fft(v)
        sgn = -1.0 if (l & 3) == 3 else 1.0          # pre-negation that rotates lane par=1's second DFT by 8
        for k1 in range(16):
            u[k1] *= cw(l * k1, 512) * sgn
        U[l] = u
    for k1 in range(16):
        sm.access(f"A.sts k1={k1}", {l: 8 * (34 * k1 + l) for l in lanes}, 8, True,
                  {l: [U[l][k1].real, U[l][k1].imag] for l in lanes})
    # ---- pass B part 1
    R = {l: np.zeros(16, complex) for l in lanes}
    for m in range(16):
        got = sm.access(f"B.lds m={m}", {l: 8 * (34 * (l >> 1) + 2 * m + (l & 1)) for l in lanes}, 8, False)
        for l in lanes:
            R[l][m] = got[l][0] + 1j * got[l][1]
    for l in lanes:
        R[l] = np.fft.fft(R[l])
        # par=0: R = E[j].  par=1: inputs n0 = 2m+1 with odd m (n0 = 3 mod 4) were negated -> R[i] = O[(i+8) mod 16]
    # ---- pair exchange by shuffle: both lanes send register 8+k, receive partner's
    Z = {l: np.zeros(32, complex) for l in lanes}      # Z[k0] for k = k1 + 16 k0 (only own half filled)
    for l in lanes:
        par = l & 1
        for k in range(8):
            own = R[l][k]
            recv = R[l ^ 1][8 + k]
            if par == 0:
                T = cw(k, 32) * recv                   # X[k] = E[k] + W32^k O[k]
                Z[l][k] = own + T
                Z[l][k + 16] = own - T
            else:
                T = cw(8 + k, 32) * own                # own = O[8+k], recv = E[8+k]
                Z[l][8 + k] = recv + T
                Z[l][24 + k] = recv - T
    # check Z against the FFT of the packed frame
    zc = (x[0::2] * (0.5 * win[0::2] if win is not None else 1.0)) + 1j * (x[1::2] * (0.5 * win[1::2] if win is not None else 1.0))
    Zref = np.fft.fft(zc)
    err = 0.0
    for l in lanes:
        k1, par = l >> 1, l & 1
        for k0 in (list(range(8)) + list(range(16, 24)) if par == 0 else list(range(8, 16)) + list(range(24, 32))):
            err = max(err, abs(Z[l][k0] - Zref[k1 + 16 * k0]))
    # ---- split: lane (k1, par) owns pairs k0 in its lower set (k < 256), publishes its upper set
    # lower set: par 0 -> k0 0..7, par 1 -> k0 8..15; upper: par 0 -> 16..23, par 1 -> 24..31
    for i in range(8):
        sm.access(f"Z.sts i={i}", {l: 8 * (l + 32 * i) for l in lanes}, 8, True,
                  {l: [Z[l][(16 if (l & 1) == 0 else 24) + i].real, Z[l][(16 if (l & 1) == 0 else 24) + i].imag] for l in lanes})
    X = np.full(M + 1, np.nan, complex)
    scale = 1.0 if win is not None else 0.5
    pw = lambda k: complex(-np.sin(2 * np.pi * k / N), -np.cos(2 * np.pi * k / N))
    for i in range(8):
        addrs = {}
        for l in lanes:
            k1, par = l >> 1, l & 1
            k0 = (0 if par == 0 else 8) + i
            # partner of k = k1 + 16 k0: k1 != 0 -> (16 - k1, k0' = 31 - k0); k1 == 0 -> (0, k0' = 32 - k0)
            if k1 != 0:
                kp1, kp0 = 16 - k1, 31 - k0
            else:
                kp1, kp0 = 0, 32 - k0
            if kp0 >= 32:                                  # k = 0: pairs with itself
                addrs[l] = 0
                continue
            ppar = 0 if (kp0 < 24) else 1                  # upper sets: 16..23 (par 0), 24..31 (par 1)
            pl = 2 * kp1 + ppar
            addrs[l] = 8 * (pl + 32 * (kp0 - (16 if ppar == 0 else 24)))
        got = sm.access(f"P.lds i={i}", addrs, 8, False)
        for l in lanes:
            k1, par = l >> 1, l & 1
            k0 = (0 if par == 0 else 8) + i
            k = k1 + 16 * k0
            A_ = Z[l][k0]
            B_ = A_ if k == 0 else got[l][0] + 1j * got[l][1]
            S = A_ + np.conj(B_)
            D = A_ - np.conj(B_)
            Q = D * pw(k)
            X[k] = (S + Q) * scale
            X[M - k] = np.conj(S - Q) * scale
    X[256] = np.conj(Zref[256]) * 2 * scale               # lane (k1=0, par=0) holds Z[256] = k0 16 (published set)
    return X, sm.log, err


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    x = rng.standard_normal(N)
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / N)
    for win in (None, w):
        X, log, err = run_frame(x, win)
        ref = np.fft.rfft(x * (win if win is not None else 1.0))
        print("Z err", err, " X max rel err", np.nanmax(np.abs(X - ref)) / np.abs(ref).max(), "nan bins", int(np.isnan(X).sum()))
    by = {}
    for name, wf, width, st in log:
        k = name.split()[0]
        by.setdefault(k, [0, 0]); by[k][0] += wf; by[k][1] += 1
    for k, (wf, n) in by.items():
        print(f"{k:8s} instr {n:3d} wavefronts {wf:4d}")
