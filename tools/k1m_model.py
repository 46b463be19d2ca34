#!/usr/bin/env python3
"""Numpy model of the team-per-frame kernels (fft_team.cuh): N = 32 T real points on T = 16 R lanes, R in {2, 4, 8}.

    M = N/2 = 16 T,  n = T n1 + n0,  k = k1 + 16 k0,  k0 = j + 16 q
    A: lane t = n0: radix-16 over n1, twiddle W_M^{n0 k1} (x W_R^{(n0 % R)(n0 / R)}: pre-rotation), publish (k1, n0)
    B: lane t = R k1 + r gathers n0 = R m + r, radix-16 -> register i holds E_r[(i + c_r) % 16], c_r = (16/R) r;
       all-to-all inside the R-lane group by shfl.idx with ROTATION: from lane (r + x) % R the registers
       (16 - (16/R) x + d) % 16 are the j = c_r + d this lane finishes -- the same register index in every lane;
       twiddle W_T^{r' j}, R-point DFT over x, phase W_R^{r q}  ->  Z[k1 + 16 (j + 16 q)]
    split: keep q < R/2, publish q >= R/2 at [slot][t]; partner of kept index i is slot 7 - i of lane T - 1 - t + R
Checks the result against numpy.fft.rfft and the bank behaviour of every shared access (64-bit, per half warp).
"""
import sys

import numpy as np


def cw(num, den):
    return np.exp(-2j * np.pi * (num % den) / den)


def half_warp_conflicts(addr_by_lane):
    """64-bit accesses: a half warp is one wavefront when its 16 element indices are distinct mod 16."""
    worst = 1
    lanes = sorted(addr_by_lane)
    for base in range(0, len(lanes), 16):
        grp = [addr_by_lane[t] for t in lanes[base:base + 16]]
        cnt = {}
        for a in set(grp):
            cnt[a % 16] = cnt.get(a % 16, 0) + 1
        worst = max(worst, max(cnt.values()))
    return worst


def run_frame(x, win, R, stats):
    T, Q = 16 * R, 16 // R
    M, N = 16 * T, 32 * T
    S = T + R
    w = 0.5 * win if win is not None else np.full(N, 0.5)
    z = x[0::2] * w[0::2] + 1j * x[1::2] * w[1::2]
    ex = np.zeros(16 * S, complex)
    # ---- A
    for t in range(T):
        u = np.fft.fft(z[t::T][:16])
        for k1 in range(16):
            ex[S * k1 + t] = u[k1] * cw(t * k1, M) * cw((t % R) * (t // R), R)
    for k1 in range(16):
        stats["A.st"] = max(stats.get("A.st", 1), half_warp_conflicts({t: S * k1 + t for t in range(T)}))
    # ---- B: gather + radix-16
    E = {}
    for t in range(T):
        k1, r = divmod(t, R)
        E[t] = np.fft.fft(np.array([ex[S * k1 + R * m + r] for m in range(16)]))
    for m in range(16):
        stats["B.ld"] = max(stats.get("B.ld", 1),
                            half_warp_conflicts({t: S * (t // R) + R * m + (t % R) for t in range(T)}))
    # rotation check: register i of lane r holds E_r[(i + c_r) % 16]
    Z = np.zeros(M, complex)
    Zr = {}
    for t in range(T):
        k1, r = divmod(t, R)
        c = Q * r
        zr = np.zeros((Q, R), complex)
        for d in range(Q):
            j = c + d
            g = np.zeros(R, complex)
            for xx in range(R):
                rp = (r + xx) % R
                src = (t & ~(R - 1)) | rp
                reg = (16 - Q * xx + d) % 16
                g[xx] = E[src][reg] * cw(rp * j, T)
            D = np.fft.fft(g)
            for q in range(R):
                zr[d, q] = D[q] * cw(r * q, R)
                Z[k1 + 16 * (j + 16 * q)] = zr[d, q]
        Zr[t] = zr
    Zref = np.fft.fft(z)
    err = np.abs(Z - Zref).max() / np.abs(Zref).max()
    assert err < 1e-12, ("packed FFT", err)
    # ---- split
    zx = np.zeros(9 * T, complex)
    for t in range(T):
        for q in range(R // 2, R):
            for d in range(Q):
                zx[((q - R // 2) * Q + d) * T + t] = Zr[t][d, q]
    X = np.zeros(M + 1, complex)
    for t in range(T):
        k1, r = divmod(t, R)
        c = Q * r
        base = T - 1 - t + R if k1 else (R - 1 - r) + T
        for q in range(R // 2):
            for d in range(Q):
                i = q * Q + d
                A = Zr[t][d, q]
                B = zx[(7 - i) * T + base]
                if k1 == 0 and d == 0:
                    if r > 0:
                        B = zx[((R // 2 - 1 - q) * Q) * T + (R - r)]
                    elif q == 0:
                        B = A
                    else:
                        B = Zr[t][0, R - q]
                k = k1 + 16 * (c + d) + 256 * q
                assert abs(B - Zref[(M - k) % M]) < 1e-9 * np.abs(Zref).max(), (t, q, d)
                Sx = A + np.conj(B)
                Dx = A - np.conj(B)
                pw_lane = -1j * cw(k1 + 16 * c, N)
                Qx = Dx * cw(d + 16 * q, 32 * R) * pw_lane
                X[k] = Sx + Qx
                X[M - k] = np.conj(Sx - Qx)
    X[M // 2] = 2 * np.conj(Zr[0][0, R // 2])
    for i in range(8):
        stats["S.ld"] = max(stats.get("S.ld", 1), half_warp_conflicts(
            {t: (7 - i) * T + (T - 1 - t + R if t >= R else (R - 1 - t) + T) for t in range(T)}))
    ref = np.fft.rfft(x * (win if win is not None else 1.0))
    err = np.abs(X - ref).max() / np.abs(ref).max()
    assert err < 1e-12, ("spectrum", err)
    return err


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for R in (2, 4, 8):
        N = 512 * R
        stats = {}
        for win in (None, np.hanning(N + 1)[:N]):
            e = run_frame(rng.standard_normal(N), win, R, stats)
        print(f"R={R} N={N}: max rel err {e:.2e}; worst half-warp multiplicity {stats}")
    sys.exit(0)
