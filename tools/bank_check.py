#!/usr/bin/env python3
"""Shared-memory wavefront model for the K1 Stockham exchanges (design tool).

For every warp-wide LDS/STS the batched FFT kernel issues, count the number of
128-byte wavefronts the access needs (1 per half-warp for 64-bit, 1 per
quarter-warp for 128-bit when conflict-free) under a candidate swizzle.  Used
to choose the swizzle in kopek_b200/csrc/fft_batched.cuh; not part of the
product or the tests.
"""
import itertools
import sys


def wavefronts(addrs_bytes, width):
    """addrs_bytes: 32 lane byte addresses; width: 8 or 16 bytes per lane."""
    lanes_per_phase = 128 // width
    total = 0
    for p in range(0, 32, lanes_per_phase):
        lanes = addrs_bytes[p:p + lanes_per_phase]
        # bank -> set of distinct 4-byte words requested
        per_bank = {}
        for a in lanes:
            for w in range(width // 4):
                word = a // 4 + w
                per_bank.setdefault(word % 32, set()).add(word)
        total += max(len(s) for s in per_bank.values())
    return total


def stockham_accesses(M, T, radices, jmap, frames_per_warp):
    """Yield (name, [32 lane complex indices]) for every warp instruction."""
    R = M // T
    Ns = 1
    for s, r in enumerate(radices):
        nb = R // r  # butterflies per thread
        if s > 0:  # read side (stage 0 reads global)
            for i in range(nb):
                for t in range(r):
                    idx = []
                    for lane in range(32):
                        f, tid = divmod(lane, T) if T < 32 else (0, lane)
                        j = jmap(tid, i, nb, T)
                        idx.append((f, j + t * (M // r)))
                    yield (f"R{s} i{i} t{t}", idx)
        for i in range(nb):
            for k in range(r):
                idx = []
                for lane in range(32):
                    f, tid = divmod(lane, T) if T < 32 else (0, lane)
                    j = jmap(tid, i, nb, T)
                    kk = j % Ns
                    idx.append((f, (j - kk) * r + kk + k * Ns))
                yield (f"W{s} i{i} k{k}", idx)
        Ns *= r
    # post-pass reads: A = Z[k], B = Z[M-k]
    for i in range(R // 2):
        idxa, idxb = [], []
        for lane in range(32):
            f, tid = divmod(lane, T) if T < 32 else (0, lane)
            k = tid + T * i
            idxa.append((f, k))
            idxb.append((f, (M - k) % M))
        yield (f"PA i{i}", idxa)
        yield (f"PB i{i}", idxb)


def evaluate(M, T, radices, jmap, swz, frame_stride, verbose=False):
    tot = ideal = 0
    worst = []
    for name, idx in stockham_accesses(M, T, radices, jmap, max(1, 32 // T)):
        addrs = [(f * frame_stride + swz(q)) * 8 for f, q in idx]
        w = wavefronts(addrs, 8)
        tot += w
        ideal += 2
        if w > 2:
            worst.append((name, w))
    if verbose:
        for n, w in worst[:20]:
            print("   ", n, w)
    return tot, ideal


def jmap_strided(tid, i, nb, T):
    return tid + T * i


if __name__ == "__main__":
    cfgs = {
        128: (8, (8, 16)), 256: (16, (16, 16)), 512: (32, (8, 8, 8)),
        1024: (64, (4, 16, 16)), 2048: (128, (8, 16, 16)),
    }
    swzs = {
        "none": lambda q: q,
        "x3": lambda q: q ^ ((q >> 3) & 7),
        "x4": lambda q: q ^ ((q >> 4) & 15),
        "x3_15": lambda q: q ^ ((q >> 3) & 15),
        "x4x8": lambda q: q ^ ((q >> 4) & 15) ^ ((q >> 8) & 15),
        "x3x7": lambda q: q ^ ((q >> 3) & 15) ^ ((q >> 7) & 15),
        "pad16": lambda q: q + (q >> 4),
        "pad8": lambda q: q + (q >> 3),
    }
    for M, (T, rad) in cfgs.items():
        for name, swz in swzs.items():
            maxq = max(swz(q) for q in range(M)) + 1
            tot, ideal = evaluate(M, T, rad, jmap_strided, swz, maxq, verbose="-v" in sys.argv)
            print(f"M={M:5d} T={T:3d} rad={rad} swz={name:6s} wavefronts={tot:5d} ideal={ideal:5d} x{tot/ideal:.2f}")
