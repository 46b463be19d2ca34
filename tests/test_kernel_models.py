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

