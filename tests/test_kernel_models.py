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



def test_bench_static_mix_matches_the_library(built_lib):
    """bench.py quotes FP32 and instruction-issue roofs from a static instruction mix of the frame loops; the numbers
    must be those of the library that ships (cuobjdump -sass, no GPU needed)."""
    import collections
    import re
    import subprocess

    import bench
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True, check=True).stdout
    blocks = {b.split("\n", 1)[0].strip(): b for b in re.split(r"\n\s*Function : ", sass)[1:]}
    for key, mix in bench.KERNEL_MIX.items():
        names = [n for n in blocks if mix["symbol"] in n]
        assert len(names) == 1, (key, names)
        ins = [(int(m.group(1), 16), m.group(2)) for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", blocks[names[0]])]
        span = (0, 0, 0)
        for a, t in ins:
            m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a and a - int(m.group(1), 16) > span[0]:
                span = (a - int(m.group(1), 16), int(m.group(1), 16), a)
        loop = [t for a, t in ins if span[1] <= a <= span[2]]
        c = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for t in loop)
        got = {"loop": len(loop), "FFMA": c["FFMA"], "FMUL": c["FMUL"], "FADD": c["FADD"], "FADD2": c["FADD2"]}
        assert got == {k: mix[k] for k in got}, (key, got)
