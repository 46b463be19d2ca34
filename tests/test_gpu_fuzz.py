"""GPU: randomised plan geometry against the oracle -- every dispatch path of kopek_stft_exec.

Frame size, hop, window, output, bin mode, row pitch, sample-pointer offset (which selects between the vector, 64-bit and
scalar load paths and therefore between the lane-group kernels and the generic one), bulk-store flag and frame count
are drawn at random (seeded); each case is compared with the oracle's f64 restatement at the north_star tolerance."""
import numpy as np
import pytest

from util import tol

pytestmark = pytest.mark.gpu

SIZES = [256, 512, 1024, 2048, 4096]


def _cases(count, seed):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(count):
        n = int(rng.choice(SIZES))
        hop = int(rng.choice([n, n // 2, n // 4, n // 4 + 2, int(rng.integers(1, 2 * n)), 4 * int(rng.integers(1, n // 2))]))
        window = int(rng.integers(0, 3))
        output = int(rng.integers(0, 4))
        full = bool(rng.integers(0, 4) == 0)
        bins = n if full else n // 2 + 1
        pitch = int(rng.choice([0, 0, bins + int(rng.integers(0, 9)), ((bins + 3) // 4) * 4]))
        offset = int(rng.choice([0, 0, 0, 1, 2, 4]))
        frames = int(rng.choice([1, 2, 3, 7, int(rng.integers(1, 400)), int(rng.integers(900, 2400))]))
        bulk = bool(rng.integers(0, 3) == 0)
        out.append((i, n, hop, window, output, full, pitch, offset, frames, bulk))
    return out


@pytest.mark.parametrize("case", _cases(72, 20261020), ids=lambda c: f"{c[0]}-n{c[1]}-h{c[2]}-w{c[3]}-o{c[4]}")
def test_random_plan_geometry(built_lib, oracle_mod, case):
    import torch
    from kopek_b200 import StftPlan, BINS_FULL, BINS_HALF
    i, n, hop, window, output, full, pitch, offset, frames, bulk = case
    rng = np.random.default_rng(1000 + i)
    n_samples = (frames - 1) * hop + n + int(rng.integers(0, hop))          # a ragged tail that makes no extra frame
    x = (rng.standard_normal(n_samples + offset) * rng.choice([1e-3, 1.0, 300.0])).astype(np.float32)
    d = torch.from_numpy(x).cuda()
    # the bulk-store epilogue exists for n = 1024, HALF bins, f32 outputs, 16-byte rows; elsewhere the setter refuses
    bins_expected = n if full else n // 2 + 1
    bulk = bulk and n == 1024 and not full and output >= 1 and (pitch or bins_expected) % 4 == 0
    with StftPlan(n, hop, window, output, BINS_FULL if full else BINS_HALF, out_pitch=pitch, bulk_store=bulk) as plan:
        assert plan.frames(n_samples) == frames
        bins = plan.bins
        out = torch.full((frames + 1, plan.pitch), float("nan"), device="cuda",
                         dtype=torch.complex64 if output == 0 else torch.float32)
        got_frames = plan.exec_ptr(d.data_ptr() + 4 * offset, n_samples, out.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert got_frames == frames
        raw = out.cpu().numpy()
    got, pad, guard = raw[:frames, :bins], raw[:frames, bins:], raw[frames:]
    assert not np.isnan(got).any(), "unwritten output element"
    assert np.isnan(pad).all() and np.isnan(guard).all(), "wrote outside the spectra"
    ref = oracle_mod.stft_complex(x[offset:offset + n_samples], n, hop, window, precision=64, bins=bins)
    top = np.abs(ref).max(axis=1, keepdims=True)
    top[top == 0] = 1.0
    if output == 0:
        err = np.abs(got.astype(np.complex128) - ref) / top
    elif output == 1:
        err = np.abs(got.astype(np.float64) - np.abs(ref)) / top
    elif output == 2:
        err = np.abs(got.astype(np.float64) - np.abs(ref) ** 2) / top ** 2 / 2
    else:
        err = np.abs(10.0 ** (got.astype(np.float64) / 20.0) - np.abs(ref)) / top
    assert err.max() <= tol(n), (case, float(err.max()))
