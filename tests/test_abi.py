"""CPU: the C-ABI library builds, loads and exports every symbol include/kopek_cuda.h declares.
No compute call is made here (there is no GPU in the dev container)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "kopek_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"KOPEK_API\s+[^;()]*?\b(kopek_\w+)\s*\(", text)))


def test_header_symbols_all_exported(built_lib):
    declared = _declared_symbols()
    assert len(declared) >= 20
    nm = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (kopek_\w+)", nm))
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in kopek_cuda.h but not exported: {missing}"
    # nothing but the ABI leaks out of the library
    extra = [s for s in exported if s not in declared]
    assert not extra, f"exported but not declared: {extra}"


def test_python_binding_covers_header(built_lib):
    from kopek_b200 import _lib
    assert sorted(_lib.SYMBOLS) == _declared_symbols()
    lib = _lib.load()
    assert lib.kopek_version().decode().startswith("kopek_b200")


def test_rust_sys_crate_declares_every_symbol():
    """rust/kopek-cuda-sys cannot be compiled here (no rustc); keep it textually in step with the header."""
    text = open(os.path.join(ROOT, "rust", "kopek-cuda-sys", "src", "lib.rs")).read()
    declared = set(re.findall(r"pub fn (kopek_\w+)\s*\(", text))
    assert declared == set(_declared_symbols())


def test_library_has_sm100a_code_only(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_host_only_entry_points(built_lib):
    from kopek_b200 import _lib, fft
    lib = _lib.load()
    for n, want in [(0, 1), (1, 1), (2, 2), (3, 4), (1000, 1024), (1024, 1024), (1025, 2048)]:
        assert lib.kopek_next_pow2(n) == want          # usize::next_power_of_two, src/fft.rs:32
    want = open(os.path.join(GOLDEN, "show.txt")).read()
    assert fft.show_format("label", [1 + 2j, -0.5 - 0.25j, 0j, complex(-0.0, -3.14159)]) == want
    assert fft.show_format("x", [complex(float("nan"), float("-inf"))]) == "x\nNaN-infi\n"
    # truncating form of the snprintf convention
    buf = C.create_string_buffer(8)
    a = np.array([1 + 2j], dtype=np.complex128)
    need = lib.kopek_show_format(b"label", a.ctypes.data, 1, C.addressof(buf), 8)
    assert need == len("label\n1.0000+2.0000i\n") and buf.value == b"label\n1"


def test_argument_validation_without_gpu(built_lib):
    from kopek_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.kopek_plan_create(C.byref(h), 1000, 1000, 0, 1, 0, 0, 0) == _lib.ERR_UNSUPPORTED
    assert b"power of two" in lib.kopek_last_error()
    assert lib.kopek_plan_create(C.byref(h), 1024, 0, 0, 1, 0, 0, 0) == _lib.ERR_INVALID
    assert lib.kopek_plan_create(C.byref(h), 1024, 512, 9, 1, 0, 0, 0) == _lib.ERR_INVALID
    assert lib.kopek_plan_create(C.byref(h), 1024, 512, 0, 1, 0, 100, 0) == _lib.ERR_INVALID   # pitch < bins
    assert lib.kopek_plan_create(None, 1024, 512, 0, 1, 0, 0, 0) == _lib.ERR_INVALID
    out = np.zeros(4, dtype=np.complex128)
    assert lib.kopek_fft_c2c_f64(out.ctypes.data, 3, out.ctypes.data, 8) == _lib.ERR_INVALID   # n_out != next_pow2
    assert lib.kopek_plan_frames(None, 10) == 0 and lib.kopek_plan_destroy(None) == _lib.OK


def test_no_cpu_fallback_without_device(built_lib):
    """On a box without a GPU the product must fail loudly, never compute on the CPU."""
    from kopek_b200 import _lib, fft
    if _lib.load().kopek_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(fft.KopekError) as ei:
        fft.fft(np.ones(8, dtype=np.complex128))
    assert ei.value.status == _lib.ERR_CUDA
    with pytest.raises(fft.KopekError):
        fft.StftPlan(1024)


def test_product_never_imports_oracle():
    """Only tests/, bench.py's CPU baseline and smoke() may touch oracle/."""
    pkg = os.path.join(ROOT, "kopek_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text and "ref_fft" not in text, f
                assert "libkopek_oracle" not in text, f


def test_c99_program_links_against_the_header_and_fails_loudly_without_a_gpu(built_lib, tmp_path):
    """The boundary is a C ABI: a plain C99 translation unit includes the header, uses every kopek_mgpu_* entry point
    and links with -lkopek_cuda.  In this container (no GPU) the first call must fail with the library's message."""
    exe = str(tmp_path / "abi_check")
    libdir = os.path.dirname(built_lib)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "abi_check.c"), "-L", libdir, "-lkopek_cuda",
                    f"-Wl,-rpath,{libdir}", "-o", exe], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    import torch
    if torch.cuda.is_available() and torch.cuda.device_count() >= 2:
        assert r.returncode == 0, r.stderr
    else:
        assert r.returncode == 1 and r.stderr.strip(), "a missing device must be an error with a message, not a fallback"
