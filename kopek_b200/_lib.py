"""ctypes binding of libkopek_cuda.so (the C ABI in include/kopek_cuda.h).

The library is the product; this module only loads it.  There is no fallback:
if the shared object is missing the import of the symbols fails loudly with the
build command to run.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libkopek_cuda.so")
if os.environ.get("KOPEK_TUNE_TAG"):      # tuning builds of the same library (tools/tune.py), never a different backend
    LIB_PATH = os.path.join(_HERE, "lib", "tune", os.environ["KOPEK_TUNE_TAG"], "libkopek_cuda.so")

OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_UNSUPPORTED = 0, -1, -2, -3, -4
WINDOW_RECT, WINDOW_HANN, WINDOW_HAMMING, WINDOW_CUSTOM = 0, 1, 2, 3
OUT_COMPLEX, OUT_MAG, OUT_POWER, OUT_DB, OUT_BANDS_LOG, OUT_BANDS_LOW = 0, 1, 2, 3, 4, 5
BAND_ROW = 10
BINS_HALF, BINS_FULL = 0, 1

# every symbol include/kopek_cuda.h declares: name -> (restype, argtypes)
_sz, _vp, _i32, _u32, _u64 = C.c_size_t, C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64
SYMBOLS = {
    "kopek_last_error": (C.c_char_p, []),
    "kopek_version": (C.c_char_p, []),
    "kopek_device_count": (_i32, []),
    "kopek_next_pow2": (_sz, [_sz]),
    "kopek_fft_c2c_f64": (_i32, [_vp, _sz, _vp, _sz]),
    "kopek_ifft_c2c_f64": (_i32, [_vp, _sz, _vp]),
    "kopek_fft_c2c_f64_batched": (_i32, [_vp, _sz, _sz, _vp]),
    "kopek_fft_c2c_f64_device": (_i32, [_vp, _sz, _sz, _vp, C.c_int, _vp]),
    "kopek_show_format": (_sz, [C.c_char_p, _vp, _sz, _vp, _sz]),
    "kopek_show": (_i32, [C.c_char_p, _vp, _sz]),
    "kopek_plan_create": (_i32, [C.POINTER(_vp), _u32, _u32, _u32, _u32, _u32, _u64, C.c_int]),
    "kopek_plan_set_window": (_i32, [_vp, _vp]),
    "kopek_plan_set_band_scale": (_i32, [_vp, C.c_float]),
    "kopek_plan_set_bulk_store": (_i32, [_vp, C.c_int]),
    "kopek_plan_destroy": (_i32, [_vp]),
    "kopek_plan_frames": (_u64, [_vp, _u64]),
    "kopek_plan_bins": (_u64, [_vp]),
    "kopek_plan_out_pitch": (_u64, [_vp]),
    "kopek_plan_out_elem_bytes": (_u64, [_vp]),
    "kopek_plan_last_launches": (_u32, [_vp]),
    "kopek_stft_exec": (_i32, [_vp, _vp, _u64, _vp, C.POINTER(_u64), _vp]),
    "kopek_stft_exec_pcm16": (_i32, [_vp, _vp, _u64, _u32, _vp, C.POINTER(_u64), _vp]),
    "kopek_stft_exec_host": (_i32, [_vp, _vp, _u64, _vp, C.POINTER(_u64)]),
    "kopek_stream_create": (_i32, [C.POINTER(_vp), _u32, _u32, _u32, C.c_float, C.c_int]),
    "kopek_stream_push": (_i32, [_vp, _vp, _u32, _vp]),
    "kopek_stream_bins": (_u64, [_vp]),
    "kopek_stream_destroy": (_i32, [_vp]),
    "kopek_synth_frames": (_i32, [_vp, _u64, _u32, _u64, C.c_float, _u32, C.c_float, C.c_float, C.c_float, _u32, _u32,
                                  C.c_float, _u64, _u64, C.c_int, _vp]),
    "kopek_pcm16_to_f32": (_i32, [_vp, _u64, _u32, _u32, C.c_double, _vp, C.c_int, _vp]),
    "kopek_pcm16_scale_f32": (_i32, [_vp, _u64, C.c_float, _vp, C.c_int, _vp]),
    "kopek_f32_downmix": (_i32, [_vp, _u64, _u32, _vp, C.c_int, _vp]),
    "kopek_detect_bpm": (_i32, [_vp, _u64, C.c_float, C.POINTER(_u32), C.POINTER(_u32), _vp, _vp, C.c_int, _vp]),
    "kopek_detect_bpm_host": (_i32, [_vp, _u64, C.c_float, C.POINTER(_u32), C.POINTER(_u32), C.c_int]),
    "kopek_device_alloc": (_i32, [C.POINTER(_vp), _sz, C.c_int]),
    "kopek_device_free": (_i32, [_vp, C.c_int]),
    "kopek_copy": (_i32, [_vp, _vp, _sz, C.c_int]),
    "kopek_ipc_export": (_i32, [_vp, _vp]),
    "kopek_ipc_open": (_i32, [_vp, C.c_int, C.POINTER(_vp)]),
    "kopek_ipc_close": (_i32, [_vp, C.c_int]),
    "kopek_fft_dist_combine": (_i32, [_vp, _u32, _u32, _u64, _vp, C.c_int, _vp]),
    "kopek_host_alloc": (_i32, [C.POINTER(_vp), _sz]),
    "kopek_host_free": (_i32, [_vp]),
    "kopek_fft_large_c2c_f32": (_i32, [_vp, _vp, _u64, C.c_int, _vp]),
    "kopek_mgpu_create": (_i32, [C.POINTER(_vp), _vp, _u32, _u32, _u32, _u32, _u32, _u32, _u64]),
    "kopek_mgpu_destroy": (_i32, [_vp]),
    "kopek_mgpu_device_count": (_u32, [_vp]),
    "kopek_mgpu_bins": (_u64, [_vp]),
    "kopek_mgpu_out_pitch": (_u64, [_vp]),
    "kopek_mgpu_out_elem_bytes": (_u64, [_vp]),
    "kopek_mgpu_set_custom_window": (_i32, [_vp, _vp]),
    "kopek_mgpu_shard": (_i32, [_vp, _u64, _u32, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64)]),
    "kopek_mgpu_exec": (_i32, [_vp, _vp, _vp, _vp, _vp]),
    "kopek_mgpu_exec_gathered": (_i32, [_vp, _vp, _vp, _vp, C.POINTER(_u64)]),
    "kopek_mgpu_gather": (_i32, [_vp, _vp, _vp, _vp, _u32]),
    "kopek_mgpu_sync": (_i32, [_vp, C.POINTER(C.c_float)]),
    "kopek_mgpu_exec_host": (_i32, [_vp, _vp, _u64, _vp, C.POINTER(_u64)]),
}
GATHER_COPY, GATHER_NCCL = 0, 1


class KopekError(RuntimeError):
    """A C-ABI call returned a negative kopek_status."""

    def __init__(self, status: int, message: str):
        super().__init__(f"kopek status {status}: {message}")
        self.status = status


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing -- build it with `python -m kopek_b200.build` "
                "(nvcc, sm_100a). kopek_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)   # AttributeError if the build is stale
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status: int) -> None:
    if status != OK:
        raise KopekError(status, load().kopek_last_error().decode(errors="replace"))
