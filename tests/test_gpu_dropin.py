"""GPU: the kopek::fft::fft drop-in (kopek_fft_c2c_f64, K2) against the oracle -- bit-exact."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from util import frame_rel_err

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


@pytest.mark.parametrize("n_in", [0, 1, 2, 3, 4, 5, 8, 16, 31, 64, 100, 256, 1000, 1024, 2048, 4096, 5000, 8192,
                                  65536, 100000])
def test_fft_bit_exact_vs_oracle(built_lib, oracle_mod, n_in):
    from kopek_b200 import fft
    rng = np.random.default_rng(1000 + n_in)
    x = rng.standard_normal(n_in) + 1j * rng.standard_normal(n_in)
    got = fft.fft(x)
    ref = oracle_mod.fft(x)
    assert got.shape == ref.shape == (oracle_mod.next_pow2(n_in),)
    # same butterflies, same twiddles, no FMA: identical bits (src/fft.rs:23-27)
    np.testing.assert_array_equal(_bits(got), _bits(ref))


def test_fft_config0_sine440(built_lib, oracle_mod):
    """BASELINE.json configs[0]: 1024-pt spectrum of the 440 Hz / 44.1 kHz Oscillator::sine."""
    from kopek_b200 import fft
    g = np.load(os.path.join(GOLDEN, "sine440_1024.npz"))
    x = oracle_mod.marshal(g["samples"])                 # f32 -> f64 -> Complex{re,0} (pong/main.rs:192-195)
    got = fft.fft(x)
    np.testing.assert_array_equal(_bits(got), _bits(oracle_mod.fft(x)))
    assert frame_rel_err(got, g["spectrum"]) < 4e-15
    mag = np.abs(got)
    assert int(mag[:513].argmax()) == 10 and abs(mag.max() - 471.9957) < 1e-3
    # beep/graph marshalling divides by i16::MAX (beep/view.rs:143)
    x2 = oracle_mod.marshal(g["samples"], 32767.0)
    np.testing.assert_array_equal(_bits(fft.fft(x2)), _bits(oracle_mod.fft(x2)))


def test_fft_large_and_special_values(built_lib, oracle_mod):
    from kopek_b200 import fft
    rng = np.random.default_rng(77)
    x = rng.standard_normal(1 << 20) + 1j * rng.standard_normal(1 << 20)
    got = fft.fft(x)
    np.testing.assert_array_equal(_bits(got), _bits(oracle_mod.fft(x)))
    assert frame_rel_err(got, np.fft.fft(x)) < 1e-14
    # impulse -> flat, DC -> bin 0 = N, non-finite input propagates like the reference
    imp = np.zeros(512, dtype=np.complex128); imp[0] = 1
    np.testing.assert_array_equal(fft.fft(imp), np.ones(512, dtype=np.complex128))
    dc = np.ones(512, dtype=np.complex128)
    y = fft.fft(dc)
    assert y[0] == 512 and np.abs(y[1:]).max() < 1e-12
    bad = np.ones(16, dtype=np.complex128); bad[3] = complex(np.inf, 0)
    np.testing.assert_array_equal(np.isnan(fft.fft(bad)), np.isnan(oracle_mod.fft(bad)))


def test_fft_batched_and_device(built_lib, oracle_mod):
    import torch
    from kopek_b200 import _lib, fft
    rng = np.random.default_rng(5)
    for n_in in (7, 256, 1000):
        x = rng.standard_normal((33, n_in)) + 1j * rng.standard_normal((33, n_in))
        got = fft.fft_batched(x)
        ref = np.stack([oracle_mod.fft(r) for r in x])
        np.testing.assert_array_equal(_bits(got), _bits(ref))
    # device-resident entry point on torch's current stream
    x = rng.standard_normal((1000, 256)) + 1j * rng.standard_normal((1000, 256))
    d_in = torch.from_numpy(x).cuda()
    d_out = torch.empty((1000, 256), dtype=torch.complex128, device="cuda")
    _lib.check(_lib.load().kopek_fft_c2c_f64_device(d_in.data_ptr(), 256, 1000, d_out.data_ptr(), 0,
                                                    torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = np.stack([oracle_mod.fft(r) for r in x])
    np.testing.assert_array_equal(_bits(d_out.cpu().numpy()), _bits(ref))


def test_show_prints_reference_format(built_lib, capfd):
    from kopek_b200 import fft
    fft.show("label", [1 + 2j, -0.5 - 0.25j, 0j, complex(-0.0, -3.14159)])
    out = capfd.readouterr().out
    assert out == open(os.path.join(GOLDEN, "show.txt")).read()


def test_cpp_mirror_harness(built_lib, oracle_mod, tmp_path):
    """kopek::fft::fft / kopek::fft::show C++ mirror (kopek_b200/host/kopek_fft.hpp): the call sequences of the
    Rust shim through the same C ABI, bit-compared with the oracle inside the harness."""
    import subprocess
    from conftest import ROOT
    exe = str(tmp_path / "test_dropin")
    libdir = os.path.dirname(built_lib)
    odir = os.path.join(ROOT, "oracle", "_build")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_dropin.cpp"),
                    "-L", libdir, "-lkopek_cuda", "-L", odir, "-lkopek_oracle",
                    f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{odir}"], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "DROPIN_OK" in r.stdout and "peak bin 10" in r.stdout
    assert "spectrum\n" in r.stdout          # show() printed the label line


def test_fft_and_plans_are_callable_from_many_threads(built_lib, oracle_mod):
    """kopek::fft::fft is a pure function called from whichever thread owns the UI (egui update, bevy system:
    SURVEY.md 8b); the drop-in and independent plans must give the same bits when eight threads hammer them at once."""
    import threading

    import torch
    from kopek_b200 import fft, StftPlan, OUT_MAG, WINDOW_HANN
    rng = np.random.default_rng(5)
    sizes = [1024, 1000, 4096, 257, 1024, 8192, 1024, 64]
    xs = [rng.standard_normal(n) + 1j * rng.standard_normal(n) for n in sizes]
    refs = [oracle_mod.fft(x) for x in xs]
    frames = rng.standard_normal(64 * 1024).astype(np.float32)
    with StftPlan(1024, 1024, WINDOW_HANN, OUT_MAG) as p0:
        plan_ref = p0.exec_host(frames)
    errors = []

    def worker(i):
        try:
            torch.cuda.set_device(0)
            with StftPlan(1024, 1024, WINDOW_HANN, OUT_MAG) as plan:
                for _ in range(25):
                    got = fft.fft(xs[i])
                    if not np.array_equal(_bits(got), _bits(refs[i])):
                        errors.append(f"thread {i}: fft() bits differ")
                        return
                    if not np.array_equal(plan.exec_host(frames), plan_ref):
                        errors.append(f"thread {i}: plan result differs")
                        return
        except Exception as exc:                                    # noqa: BLE001
            errors.append(f"thread {i}: {exc!r}")

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(len(sizes))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors


@pytest.mark.parametrize("n", [1, 2, 8, 1024, 4096, 1 << 16, 1 << 20])
def test_inverse_round_trip(built_lib, oracle_mod, n):
    """fft -> ifft is the identity: a size-independent property of the forward path, and the inverse itself equals
    the composition conj(oracle.fft(conj(X))) / n bit for bit."""
    from kopek_b200 import fft, KopekError
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    X = fft.fft(x)
    back = fft.ifft(X)
    assert np.abs(back - x).max() <= 1e-13 * max(1.0, np.log2(n)) * np.abs(x).max()
    ref = np.conj(oracle_mod.fft(np.conj(X))) / n
    np.testing.assert_array_equal(_bits(back), _bits(ref))
    if n == 8:
        with pytest.raises(KopekError):
            fft.ifft(X[:6])                       # an inverse cannot zero-pad


def test_latency_path_sees_fresh_data_every_call(built_lib, oracle_mod):
    """The host-buffer path keeps its staging buffer mapped into the device and polls a completion token instead of
    synchronising the stream: consecutive calls of every size class (single CTA with token, batched, chunked + level
    passes) with new data each time must each return the transform of THAT call's input, bit for bit."""
    from kopek_b200 import fft
    rng = np.random.default_rng(99)
    sizes = [1024, 1000, 256, 4096, 1, 2, 3, 8, 100, 8192, 1024, 5000, 65536, 1024, 16, 4095, 2048, 1024]
    for rep in range(3):
        for n in sizes:
            x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex128)
            got = fft.fft(x)
            ref = oracle_mod.fft(x)
            assert np.array_equal(got.view(np.uint64), ref.view(np.uint64)), (rep, n)
