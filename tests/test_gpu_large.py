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


def test_large_fft_same_length_on_two_streams(built_lib):
    """Round-1 advice: two calls of the same length on different streams shared one pass-1 -> pass-2 workspace and
    could overwrite each other's intermediate.  The workspace is now per call (stream-ordered pool): interleaved calls
    on two streams, many times over, give exactly the bits of the calls made one after the other."""
    import torch
    from kopek_b200 import fft
    n = 1 << 20
    g = torch.Generator(device="cuda").manual_seed(3)
    xs = [torch.randn(n, dtype=torch.complex64, device="cuda", generator=g) for _ in range(4)]
    refs = []
    st0 = torch.cuda.current_stream().cuda_stream
    for x in xs:
        o = torch.empty_like(x)
        fft.fft_large_device(x.data_ptr(), o.data_ptr(), n, 0, st0)
        torch.cuda.synchronize()
        refs.append(o)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = [torch.zeros_like(x) for x in xs]
    torch.cuda.synchronize()
    for rep in range(25):
        for i, x in enumerate(xs):
            s = s1 if i % 2 == 0 else s2
            fft.fft_large_device(x.data_ptr(), outs[i].data_ptr(), n, 0, s.cuda_stream)
    torch.cuda.synchronize()
    for o, r in zip(outs, refs):
        assert torch.equal(o.view(torch.float32), r.view(torch.float32))


def test_drop_in_many_lengths_keep_their_twiddle_tables(built_lib, oracle_mod):
    """Round-1 advice: the twiddle tables of the f64 drop-in were an LRU of eight that freed the oldest table under
    callers that might still launch on it.  They are kept for the life of the process now: twelve distinct lengths,
    visited twice in different orders, stay bit-exact."""
    from kopek_b200 import fft
    rng = np.random.default_rng(12)
    lengths = [2, 8, 16, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384]
    xs = {n: rng.standard_normal(n) + 1j * rng.standard_normal(n) for n in lengths}
    for order in (lengths, lengths[::-1], lengths[::2] + lengths[1::2]):
        for n in order:
            got = fft.fft(xs[n])
            assert np.array_equal(got.view(np.uint64), oracle_mod.fft(xs[n]).view(np.uint64)), n
