"""Build libkopek_cuda.so in-tree with nvcc for sm_100a.

    python -m kopek_b200.build [--force] [--verbose]

Explicit nvcc commands (no torch, no cmake): the five K1 frame sizes compile as
separate translation units in parallel, then everything links into
kopek_b200/lib/libkopek_cuda.so with the CUDA runtime linked statically so the
library has no dependency beyond libcuda.  The .so is git-ignored but ships to
the GPU box with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "lib", "obj")
LIB = os.path.join(HERE, "lib", "libkopek_cuda.so")

K1_SIZES = (256, 512, 1024, 2048, 4096)
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr",
          "-I", os.path.join(ROOT, "include")]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _sources():
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))]
    srcs.append(os.path.join(ROOT, "include", "kopek_cuda.h"))
    return srcs


def _fingerprint(extra) -> str:
    """Content hash of every source + the flags: the library is rebuilt exactly when one of them changed.  (Modification
    times do not survive the snapshot that ships the tree to the GPU box; the stamp file next to the .so does.)"""
    h = hashlib.sha256()
    for path in _sources():
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as f:
            h.update(f.read())
    h.update(" ".join([*ARCH, *COMMON[:-1], *extra]).encode())
    return h.hexdigest()


def _stale(target: str, stamp: str) -> bool:
    if not os.path.exists(target):
        return True
    try:
        with open(target + ".stamp") as f:
            return f.read().strip() != stamp
    except OSError:
        return True


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    if verbose and (r.stdout or r.stderr):
        print(r.stdout + r.stderr, flush=True)


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False, tag: str = "", defs=()) -> str:
    """tag / defs: tuning builds only (tools/): extra -D flags, objects and library under lib/tune/<tag>/."""
    global OBJ, LIB
    if tag:
        OBJ = os.path.join(HERE, "lib", "tune", tag, "obj")
        LIB = os.path.join(HERE, "lib", "tune", tag, "libkopek_cuda.so")
    extra = list(defs)
    if os.environ.get("KOPEK_K1_PREF") is not None:   # tuning: force the K1 register prefetch on/off for every size
        extra.append("-DK1_PREF_ALL=" + ("true" if os.environ["KOPEK_K1_PREF"] == "1" else "false"))
    stamp = _fingerprint(extra)
    if not force and not ptxas_info and not _stale(LIB, stamp):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    cc = nvcc()
    if ptxas_info:
        extra = ["-Xptxas", "-v", *extra]
    jobs = []
    for n in K1_SIZES:
        obj = os.path.join(OBJ, f"k1_{n}.o")
        jobs.append((obj, [cc, *ARCH, *COMMON, *extra, f"-DKOPEK_K1_N={n}", "-c", os.path.join(CSRC, "k1_inst.cu"),
                           "-o", obj]))
    for r in (4, 8):
        obj = os.path.join(OBJ, f"k1m_{r}.o")
        jobs.append((obj, [cc, *ARCH, *COMMON, *extra, f"-DKOPEK_K1M_R={r}", "-c", os.path.join(CSRC, "k1m_inst.cu"),
                           "-o", obj]))
    for name in ("kopek_cuda", "k1w_inst", "k1w_small_inst", "fft_large", "synth", "pcm", "dist_fft", "mgpu"):
        src = os.path.join(CSRC, name + ".cu")
        if os.path.exists(src):
            obj = os.path.join(OBJ, name + ".o")
            jobs.append((obj, [cc, *ARCH, *COMMON, *extra, "-c", src, "-o", obj]))
    with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
        list(ex.map(lambda j: _run(j[1], verbose or ptxas_info), jobs))
    _run([cc, *ARCH, "-shared", "-cudart", "static", "-o", LIB, *[j[0] for j in jobs], "-ldl"], verbose)
    with open(LIB + ".stamp", "w") as f:
        f.write(stamp + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, ptxas_info="--ptxas" in sys.argv))
