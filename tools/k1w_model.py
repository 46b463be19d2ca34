#!/usr/bin/env python3
"""Numpy model of the warp-per-frame K1 kernel (N = 1024, M = 512, 8 x 8 x 8).

Emulates the 32 lanes of one warp with the exact shared-memory address
functions of kopek_b200/csrc/fft_warp1024.cuh, checks the spectrum against
numpy.fft.rfft and counts the 128-byte shared-memory wavefronts of every
warp-wide access (design tool; not part of the product or the tests).
"""
import numpy as np

M, N = 512, 1024


def pad(L):
    """float4-slot padding of the exchange buffer: one spare slot per 8."""
    return L + (L >> 3)


class Smem:
    def __init__(self, nbytes):
        self.w = np.zeros(nbytes // 4, dtype=np.float64)   # f64 payload per 4-byte word (model precision)
        self.log = []

    def access(self, name, addrs, width, store, values=None):
        """addrs: {lane: byte address}; width bytes per lane; values: {lane: list of floats}."""
        lanes_per_phase = 128 // width
        total = 0
        for p in range(0, 32, lanes_per_phase):
            per_bank = {}
            for lane in range(p, p + lanes_per_phase):
                if lane not in addrs:
                    continue
                a = addrs[lane]
                assert a % width == 0, (name, lane, a)
                for k in range(width // 4):
                    word = a // 4 + k
                    per_bank.setdefault(word % 32, set()).add(word)
            if per_bank:
                total += max(len(s) for s in per_bank.values())
        self.log.append((name, total, width, store))
        out = {}
        for lane, a in addrs.items():
            if store:
                for k, v in enumerate(values[lane]):
                    self.w[a // 4 + k] = v
            else:
                out[lane] = [self.w[a // 4 + k] for k in range(width // 4)]
        return out


def cw(num, den):
    return np.exp(-2j * np.pi * num / den)


def dft8(v):
    return np.fft.fft(v)


def run_frame(x, win=None):
    """x: 1024 real samples.  Returns (X[0..512], smem access log)."""
    sm = Smem(288 * 16)
    z4 = x.reshape(256, 4)                       # float4 view of the frame
    w4 = (0.5 * win).reshape(256, 4) if win is not None else None
    lanes = range(32)
    # ---- stage A: n1 = l>>2, a = l&3; d[e][n2]
    dA = {}
    for l in lanes:
        d = np.zeros((2, 8), complex)
        for n2 in range(8):
            v = z4[32 * n2 + l]
            if w4 is not None:
                v = v * w4[32 * n2 + l]
            d[0, n2] = v[0] + 1j * v[1]
            d[1, n2] = v[2] + 1j * v[3]
        A = np.stack([dft8(d[0]), dft8(d[1])])
        for k2 in range(1, 8):
            A[0, k2] *= cw(l * k2, 256)
            A[1, k2] *= cw(l * k2, 256) * cw(k2, 512)
        dA[l] = A
    for k2 in range(8):
        sm.access(f"A.sts k2={k2}", {l: 16 * pad(32 * k2 + l) for l in lanes}, 16, True,
                  {l: [dA[l][0, k2].real, dA[l][0, k2].imag, dA[l][1, k2].real, dA[l][1, k2].imag] for l in lanes})
    # ---- stage B: k2 = l>>2, b = l&3 (in place)
    dB = {l: np.zeros((2, 8), complex) for l in lanes}
    for n1 in range(8):
        got = sm.access(f"B.lds n1={n1}", {l: 16 * (36 * (l >> 2) + (l & 3) + 4 * n1 + (n1 >> 1)) for l in lanes}, 16, False)
        for l in lanes:
            dB[l][0, n1] = got[l][0] + 1j * got[l][1]
            dB[l][1, n1] = got[l][2] + 1j * got[l][3]
    for l in lanes:
        b = l & 3
        B = np.stack([dft8(dB[l][0]), dft8(dB[l][1])])
        for k1 in range(1, 8):
            B[0, k1] *= cw(b * k1, 32)
            B[1, k1] *= cw(b * k1, 32) * cw(k1, 64)
        dB[l] = B
    for k1 in range(8):
        sm.access(f"B.sts k1={k1}", {l: 16 * (36 * (l >> 2) + (l & 3) + 4 * k1 + (k1 >> 1)) for l in lanes}, 16, True,
                  {l: [dB[l][0, k1].real, dB[l][0, k1].imag, dB[l][1, k1].real, dB[l][1, k1].imag] for l in lanes})
    # ---- stage C: k2 = l>>2, c = l&3, k1 = 2c+f; reads slot 36k2 + 9c + 4f + s
    Z = {l: np.zeros((2, 8), complex) for l in lanes}
    for f in range(2):
        d = {l: np.zeros(8, complex) for l in lanes}
        for s in range(4):
            got = sm.access(f"C.lds f={f} s={s}", {l: 16 * (36 * (l >> 2) + 9 * (l & 3) + 4 * f + s) for l in lanes}, 16, False)
            for l in lanes:
                d[l][2 * s] = got[l][0] + 1j * got[l][1]
                d[l][2 * s + 1] = got[l][2] + 1j * got[l][3]
        for l in lanes:
            Z[l][f] = dft8(d[l])          # Z[k2 + 16c + 8f + 64k0], k0 = 0..7
    # ---- post-pass: own pairs k0 < 4; upper half (k0 >= 4) published for the partners
    for k0 in range(4, 8):
        sm.access(f"Z.sts k0={k0}", {l: 16 * (((l + 4) & 31) + 32 * (k0 - 4)) for l in lanes}, 16, True,
                  {l: [Z[l][0, k0].real, Z[l][0, k0].imag, Z[l][1, k0].real, Z[l][1, k0].imag] for l in lanes})
    X = np.full(M + 1, np.nan, complex)
    scale = 1.0 if win is not None else 0.5
    pw = lambda k: complex(-np.sin(2 * np.pi * k / N), -np.cos(2 * np.pi * k / N))
    for k0 in range(4):
        # every lane: slot of the partner lane, k0' = 7 - k0
        got = sm.access(f"P.lds k0={k0}", {l: 16 * ((((35 - l if l >= 4 else 3 - l) + 4) & 31) + 32 * (3 - k0)) for l in lanes}, 16, False)
        # lanes 0..3 (k2 = 0): the partner of the f = 0 element lives in another slot
        addr2 = {}
        for l in range(4):
            c = l
            if c == 0 and k0 == 0:
                continue
            kk = (8 - k0) if c == 0 else (7 - k0)
            addr2[l] = 16 * (((((-c) & 3) + 4) & 31) + 32 * (kk - 4))
        got2 = sm.access(f"P2.lds k0={k0}", addr2, 16, False) if addr2 else {}
        for l in lanes:
            k2, c = l >> 2, l & 3
            lo = got[l][0] + 1j * got[l][1]
            hi = got[l][2] + 1j * got[l][3]
            for f in range(2):
                k = k2 + 16 * c + 8 * f + 64 * k0
                a = Z[l][f, k0]
                if l >= 4:
                    bq = hi if f == 0 else lo
                elif f == 1:
                    bq = hi
                elif l == 0 and k0 == 0:
                    bq = a                      # k = 0: Z[M] == Z[0]
                else:
                    bq = got2[l][0] + 1j * got2[l][1]
                S = a + np.conj(bq)
                D = a - np.conj(bq)
                Q = D * pw(k)
                X[k] = (S + Q) * scale
                X[M - k] = np.conj(S - Q) * scale
    # the lone k = 256 pair: lane 0 owns Z[256] (k2=0,c=0,f=0,k0=4): X[256] = conj Z[256] (x2 scale convention)
    X[256] = np.conj(Z[0][0, 4]) * 2 * scale
    return X, sm.log


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    x = rng.standard_normal(N)
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / N)
    for win in (None, w):
        X, log = run_frame(x, win)
        ref = np.fft.rfft(x * (win if win is not None else 1.0))
        print("max rel err", np.abs(X - ref).max() / np.abs(ref).max())
    tot = 0
    by = {}
    for name, wf, width, st in log:
        key = name.split()[0]
        by.setdefault(key, [0, 0])
        by[key][0] += wf
        by[key][1] += 1
        tot += wf
    for k, (wf, n) in by.items():
        print(f"{k:8s} instr {n:3d} wavefronts {wf:4d}")
    print("total exchange wavefronts per frame:", tot)
