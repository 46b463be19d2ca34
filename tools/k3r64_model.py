#!/usr/bin/env python3
"""numpy model of k3_pass2_r64 (kopek_b200/csrc/fft_large.cu): the 4096-point row transform as TWO radix-64 stages with a
padded in-place exchange, executed thread by thread with the kernel's own address functions, checked against numpy.fft
and for shared-memory bank conflicts (64-bit accesses: a warp's 32 lanes need 256 bytes = two wavefronts at best)."""
import sys

import numpy as np

M, COLS = 4096, 4
rng = np.random.default_rng(0)
rows = rng.standard_normal((COLS, M)) + 1j * rng.standard_normal((COLS, M))      # A[k1 = 4u + c][n2]

# the tile as pass 1 leaves it in `mid`: tile-major [n2/2][k1%4][n2%2]
tile = np.zeros(M * COLS, complex)
for c in range(COLS):
    for n2 in range(M):
        tile[(n2 >> 1) * 8 + 2 * c + (n2 & 1)] = rows[c, n2]

threads = [(tid, c) for tid in range(64) for c in range(COLS)]                    # threadIdx.x = 4 tid + c


def wavefronts(addr_f2):
    """addr_f2[thread] = float2 index accessed by one instruction; returns the worst wavefront count over the warps."""
    worst = 0
    for w in range(0, len(threads), 32):
        banks = {}
        for a in addr_f2[w:w + 32]:
            banks.setdefault((a * 8 // 8) % 16, set()).add(a)      # 16 slots of 8 bytes per 128-byte wavefront
        worst = max(worst, max(len(v) for v in banks.values()))
    return worst


buf = np.zeros((M + 64) * COLS, complex)
buf[:M * COLS] = tile
regs = {}
worst = {"A read": 0, "A write": 0, "B read": 0}
# stage A: thread (tid, c) reads n2 = tid + 64 t, transforms over t, writes output k at q = 64 tid + k -> 65 tid + k
for t in range(64):
    worst["A read"] = max(worst["A read"], wavefronts([(tid >> 1) * 8 + 2 * c + (tid & 1) + 256 * t for tid, c in threads]))
for tid, c in threads:
    v = np.array([buf[(tid >> 1) * 8 + 2 * c + (tid & 1) + 256 * t] for t in range(64)])
    regs[tid, c] = np.fft.fft(v)
for k in range(64):
    worst["A write"] = max(worst["A write"], wavefronts([(65 * tid) * 4 + c + 4 * k for tid, c in threads]))
for tid, c in threads:
    for k in range(64):
        buf[(65 * tid) * 4 + c + 4 * k] = regs[tid, c][k]
# stage B: reads q = tid + 64 t -> tid + 65 t, twiddle W4096^{tid t}, transforms over t, output k2 = tid + 64 k
out = np.zeros((COLS, M), complex)
for t in range(64):
    worst["B read"] = max(worst["B read"], wavefronts([tid * 4 + c + 260 * t for tid, c in threads]))
for tid, c in threads:
    v = np.array([buf[tid * 4 + c + 260 * t] * np.exp(-2j * np.pi * tid * t / M) for t in range(64)])
    X = np.fft.fft(v)
    for k in range(64):
        out[c, tid + 64 * k] = X[k]
ref = np.fft.fft(rows, axis=1)
err = np.abs(out - ref).max() / np.abs(ref).max()
print(f"k3_pass2_r64 model: max rel err {err:.2e}; worst wavefronts per 64-bit warp access {worst} (2 = conflict free)")
sys.exit(0 if err < 1e-12 and max(worst.values()) == 2 else 1)
