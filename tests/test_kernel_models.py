"""CPU: the numpy models of the lane-group kernels' address functions (tools/*_model.py).

Each model executes the kernel's decomposition lane by lane -- which lane loads, publishes, gathers, shuffles and
finishes which element -- and asserts the result against numpy.fft.rfft plus the bank behaviour of every shared
access.  They are the design documents the CUDA kernels were written from (fft_warp256.cuh, fft_warp1024.cuh,
fft_warp1024b.cuh, fft_team.cuh); running them here keeps model and kernel comments honest without a GPU.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("model", ["k1m_model.py", "k1w1024v2_model.py", "k1w256_model.py", "k1w_model.py", "k3r64_model.py"])
def test_model_reproduces_rfft(model):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", model)], capture_output=True, text=True, timeout=300,
                       cwd=os.path.join(ROOT, "tools"))
    assert r.returncode == 0, r.stdout + r.stderr


def test_register_butterflies_on_the_host(tmp_path):
    """Dft<2|4|8|16|64>::run are __host__ __device__: the code the kernels unroll, compiled for the CPU by nvcc and
    compared with a direct f64 DFT (tests/cpp/test_dft_host.cu)."""
    import shutil
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    exe = str(tmp_path / "test_dft_host")
    subprocess.run([nvcc, "-std=c++17", "--expt-relaxed-constexpr", "-o", exe,
                    os.path.join(ROOT, "tests", "cpp", "test_dft_host.cu")], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert "radix 64" in r.stdout
