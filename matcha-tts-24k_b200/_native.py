"""ctypes binding of the C ABI declared in include/cfm_b200.h (no torch types cross it)."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcfm_b200.so")

PREC = {"bf16": 0, "fp32": 1, "fp32_tc": 2}
SOLVERS = {"euler": 0, "midpoint": 1, "heun3": 2, "rk4": 3}
FLAG_NO_GRAPH, FLAG_SIMT_GEMM, FLAG_UNFUSED_STATS, FLAG_SIMT_ATTN = 1, 2, 4, 8
EXPORTS = ["cfm_create", "cfm_destroy", "cfm_last_error", "cfm_load_weights", "cfm_plan", "cfm_solve", "cfm_solve_host",
           "cfm_estimator", "cfm_plan_info", "cfm_debug_read", "cfm_debug_gemm", "cfm_debug_stop_after", "cfm_debug_gemm_profile", "cfm_debug_attn_profile", "cfm_set_speakers", "cfm_set_lanes", "cfm_set_option",
           "cfm_solve_host_spks", "cfm_estimator_t", "cfm_debug_timeline", "cfm_debug_ff_profile", "cfm_debug_rowln_profile", "cfm_solve_host_indexed", "cfm_synchronize",
           "cfm_front_durations", "cfm_front_expand", "cfm_denormalize"]


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("in_channels", "out_channels", "channels", "n_heads", "head_dim", "n_blocks",
                                         "n_mid_blocks", "precision", "device", "flags")]


class WeightDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("ndim", C.c_int32), ("shape", C.c_int64 * 4)]


_lib = None


def load_library(build_if_missing: bool = False) -> C.CDLL:
    """Loads libcfm_b200.so from the package directory.  There is no fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        from . import build as _build
        _build.build()
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -m matcha_tts_24k_b200.build` (needs nvcc). "
                           "This package has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64p = C.c_void_p, C.c_int32, C.POINTER(C.c_int64)
    lib.cfm_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    lib.cfm_destroy.argtypes = [vp]
    lib.cfm_destroy.restype = None
    lib.cfm_last_error.argtypes = [vp]
    lib.cfm_last_error.restype = C.c_char_p
    lib.cfm_load_weights.argtypes = [vp, C.POINTER(WeightDesc), i32]
    lib.cfm_plan.argtypes = [vp, C.POINTER(i32), i32, i32, C.POINTER(C.c_float), i32, i32]
    lib.cfm_solve.argtypes = [vp, vp, vp, vp, vp]
    lib.cfm_solve_host.argtypes = [vp, vp, vp, vp]
    lib.cfm_estimator.argtypes = [vp, vp, vp, C.c_float, vp, vp]
    lib.cfm_estimator_t.argtypes = [vp, vp, vp, C.POINTER(C.c_float), i32, vp, vp]
    lib.cfm_solve_host_spks.argtypes = [vp, vp, vp, vp, vp]
    lib.cfm_debug_timeline.argtypes = [vp, vp, vp, vp, C.c_char_p, C.c_int64, vp]
    lib.cfm_plan_info.argtypes = [vp, i64p, i64p, i64p, i64p, i64p]
    lib.cfm_debug_read.argtypes = [vp, C.c_char_p, vp, C.c_int64, i64p, i64p]
    lib.cfm_debug_gemm_profile.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, C.POINTER(i32), i32, vp, vp]
    lib.cfm_set_speakers.argtypes = [vp, vp]
    lib.cfm_set_lanes.argtypes = [vp, i32, i32]
    lib.cfm_set_option.argtypes = [vp, C.c_char_p, i32]
    lib.cfm_debug_attn_profile.argtypes = [vp, vp]
    lib.cfm_debug_ff_profile.argtypes = [vp, vp]
    lib.cfm_debug_rowln_profile.argtypes = [vp, vp]
    lib.cfm_solve_host_indexed.argtypes = [vp, vp, vp, vp, vp, C.POINTER(i32), i32]
    lib.cfm_synchronize.argtypes = [vp]
    lib.cfm_front_durations.argtypes = [vp, vp, i32, i32, vp, vp, vp]
    lib.cfm_front_expand.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp, vp, vp]
    lib.cfm_denormalize.argtypes = [vp, vp, i32, i32, i32, C.c_float, C.c_float, vp, vp]
    lib.cfm_debug_stop_after.argtypes = [vp, C.c_int64]
    lib.cfm_debug_gemm.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, C.POINTER(i32), i32, vp]
    for name in EXPORTS:
        if name not in ("cfm_destroy", "cfm_last_error"):
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


class NativeError(RuntimeError):
    pass


def check(lib, handle, status: int):
    if status == 0:
        return
    msg = lib.cfm_last_error(handle)
    msg = msg.decode() if msg else "unknown error"
    if status == -1:
        raise ValueError(msg)
    raise NativeError(f"libcfm_b200 error {status}: {msg}")
