"""Shared parity helpers (SURVEY.md 8c parity definition)."""
import math

import numpy as np


def frame_rel_err(got, ref):
    """max_k |got - ref| / max_k |ref| per frame (rows); returns the worst frame."""
    got = np.atleast_2d(got)
    ref = np.atleast_2d(ref)
    den = np.abs(ref).max(axis=1)
    den[den == 0] = 1.0
    return float((np.abs(got - ref).max(axis=1) / den).max())


def tol(n: int) -> float:
    """north_star tolerance: 1e-5 * log2(N) per bin, relative to the frame's largest bin."""
    return 1e-5 * math.log2(n)


def peak_bins(mag):
    """argmax over bins per frame; np.argmax breaks ties toward the lower index."""
    return np.argmax(np.atleast_2d(mag), axis=1)


def peaks_match(mag_got, mag_ref, n):
    """Exact peak-bin equality, except where the oracle's top two bins are closer
    than the parity tolerance (a genuine near-tie: either index is correct)."""
    mag_got = np.atleast_2d(mag_got)
    mag_ref = np.atleast_2d(mag_ref)
    pg, pr = peak_bins(mag_got), peak_bins(mag_ref)
    bad = []
    for f in np.nonzero(pg != pr)[0]:
        top = mag_ref[f, pr[f]]
        if abs(mag_ref[f, pg[f]] - top) > 2 * tol(n) * top:
            bad.append(int(f))
    return bad
