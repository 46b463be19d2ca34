"""GPU: the four-step large transform (K3, kopek_fft_large_c2c_f32) against the oracle / numpy."""
import numpy as np
import pytest

from util import frame_rel_err, tol

pytestmark = pytest.mark.gpu


def _run(x):
    import torch
    from kopek_b200 import fft
    n = x.size
    d_in = torch.from_numpy(np.ascontiguousarray(x, dtype=np.complex64)).cuda()
    d_out = torch.full((n,), float("nan"), dtype=torch.complex64, device="cuda")
    fft.fft_large_device(d_in.data_ptr(), d_out.data_ptr(), n, 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return d_out.cpu().numpy().astype(np.complex128)


@pytest.mark.parametrize("log_n", [16, 17, 18, 19, 20, 21, 22, 23, 24])
def test_large_fft_vs_numpy(built_lib, log_n):
    n = 1 << log_n
    rng = np.random.default_rng(log_n)
    # two tones + noise, seeded
    t = np.arange(n)
    x = (np.exp(2j * np.pi * 0.0123 * t) + 0.5 * np.exp(-2j * np.pi * 0.31 * t)
         + 0.1 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    got = _run(x)
    ref = np.fft.fft(x.astype(np.complex128))
    assert frame_rel_err(got, ref) <= tol(n)
    assert int(np.abs(got).argmax()) == int(np.abs(ref).argmax())


def test_large_fft_vs_oracle_and_properties(built_lib, oracle_mod):
    n = 1 << 18
    x = (np.random.default_rng(7).standard_normal(n) + 1j * np.random.default_rng(8).standard_normal(n)).astype(np.complex64)
    got = _run(x)
    ref64 = oracle_mod.fft(x.astype(np.complex128))           # the reference recursion itself, f64
    ref32 = oracle_mod.fft(x, np.complex64).astype(np.complex128)
    assert frame_rel_err(got, ref64) <= tol(n) and frame_rel_err(got, ref32) <= tol(n)
    # impulse at position p -> pure phase ramp of magnitude 1; DC -> n at bin 0
    imp = np.zeros(n, dtype=np.complex64); imp[3] = 1
    y = _run(imp)
    np.testing.assert_allclose(np.abs(y), 1.0, atol=2e-6)
    np.testing.assert_allclose(y, np.exp(-2j * np.pi * 3 * np.arange(n) / n), atol=1e-5)
    dc = _run(np.ones(n, dtype=np.complex64))
    assert abs(dc[0] - n) < 1e-3 * n and np.abs(dc[1:]).max() < 1e-4 * n


def test_large_fft_argument_checks(built_lib):
    import torch
    from kopek_b200 import fft, KopekError
    d = torch.zeros(1 << 16, dtype=torch.complex64, device="cuda")
    o = torch.zeros_like(d)
    with pytest.raises(KopekError):
        fft.fft_large_device(d.data_ptr(), d.data_ptr(), 1 << 16)          # in place
    with pytest.raises(KopekError):
        fft.fft_large_device(d.data_ptr(), o.data_ptr(), 1 << 15)          # too small
    with pytest.raises(KopekError):
        fft.fft_large_device(d.data_ptr(), o.data_ptr(), 100000)           # not a power of two
