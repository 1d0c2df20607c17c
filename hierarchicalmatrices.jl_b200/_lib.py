"""ctypes binding of the C-ABI library (include/hmb200.h -> lib/libhmb200.so).

There is no fallback: if the library has not been built this module raises, and
every compute call fails loudly when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libhmb200.so")
CSRC = os.path.join(_HERE, "csrc")

_dp = C.POINTER(C.c_double)
_i64 = C.c_int64
_i32 = C.c_int32
_vp = C.c_void_p

HM_OK = 0
STATUS_NAMES = {
    0: "HM_OK", 1: "HM_ERR_INVALID", 2: "HM_ERR_NULL", 3: "HM_ERR_SHAPE", 4: "HM_ERR_RANGE",
    5: "HM_ERR_STATE", 6: "HM_ERR_NOMEM", 7: "HM_ERR_CUDA", 8: "HM_ERR_UNSUPPORTED",
    9: "HM_ERR_REFERENCE",
}


class HmError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {msg}")
        self.status = status


class Stats(C.Structure):
    _fields_ = [(n, _i64) for n in (
        "nrows", "ncols", "n_dense", "n_lowrank", "n_bary2d", "dense_words", "lowrank_words",
        "core_words", "algorithmic_bytes", "row_begin", "row_end", "part_words", "stored_bytes",
        "v_stream_bytes", "u_stream_bytes", "partial_bytes", "n_stage1_items", "n_stage2_blocks",
        "n_stage3_items", "n_stage3_rounds", "part_algorithmic_bytes", "part_v_words", "part_core_words",
        "part_u_words", "part_dense_words")]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class TreeLeaf(C.Structure):
    _fields_ = [("kind", _i32), ("rank", _i32), ("row0", _i64), ("col0", _i64), ("m", _i64), ("n", _i64),
                ("xi0", _i64), ("yj0", _i64), ("a", C.c_double), ("b", C.c_double), ("c", C.c_double),
                ("d", C.c_double)]


# hm_kernel_fn: out[i] = f(x[i], y[i]) for i < n
KERNEL_FN = C.CFUNCTYPE(None, _dp, _dp, _i64, _dp, _vp)

# name -> (restype, argtypes); must list every symbol include/hmb200.h declares
SIGNATURES = {
    "hm_last_error": (C.c_char_p, []),
    "hm_version": (_i32, []),
    "hm_blockrank_f64": (_i32, []),
    "hm_blocksize_f64": (_i32, []),
    "hm_builder_create": (_i32, [C.POINTER(_vp), _i64, _i64, _i32, _i32]),
    "hm_builder_destroy": (_i32, [_vp]),
    "hm_builder_add_dense": (_i32, [_vp, _dp, _i64, _i64, _i64, _i64, _i64]),
    "hm_builder_add_lowrank": (_i32, [_vp, _dp, _i64, _dp, _dp, _i64, _i64, _i64, _i64, _i64, _i64]),
    "hm_builder_add_bary2d": (_i32, [_vp, _dp, _i64, _dp, _i64, _dp, _i64, _i64, _i64, _i64, _i64, _i64]),
    "hm_builder_add_evenbary": (_i32, [_vp, _dp, _i64, _dp, _i64, _i64, _i64, _i64, _i64, _i64, _i32]),
    "hm_builder_layout_stats": (_i32, [_vp, _i32, _i32, C.POINTER(Stats)]),
    "hm_plan_finalize": (_i32, [_vp, C.POINTER(_i32), _i32, C.POINTER(_vp)]),
    "hm_plan_finalize_part": (_i32, [_vp, _i32, _i32, C.POINTER(_vp)]),
    "hm_plan_destroy": (_i32, [_vp]),
    "hm_plan_stats": (_i32, [_vp, C.POINTER(Stats)]),
    "hm_assemble_kernel": (_i32, [_dp, _i64, _dp, _i64, C.c_double, C.c_double, C.c_double, C.c_double,
                                  _i32, _i32, _i32, _i32, C.POINTER(_vp)]),
    "hm_assemble_kernel_fn": (_i32, [_dp, _i64, _dp, _i64, C.c_double, C.c_double, C.c_double, C.c_double,
                                     KERNEL_FN, _vp, _i32, _i32, _i32, C.POINTER(_vp)]),
    "hm_plan_form": (_i32, [C.c_void_p, C.POINTER(C.c_int32)]),
    "hm_assemble_kernel_free": (_i32, [_dp, _i64, _dp, _i64, C.c_double, C.c_double, C.c_double, C.c_double,
                                       _i32, _i32, _i32, _i32, C.POINTER(_vp)]),
    "hm_assemble_kernel_stats": (_i32, [_dp, _i64, _dp, _i64, C.c_double, C.c_double, C.c_double,
                                        C.c_double, _i32, _i32, C.POINTER(Stats)]),
    "hm_kernel_tree_leaves": (_i32, [_dp, _i64, _dp, _i64, C.c_double, C.c_double, C.c_double, C.c_double,
                                     C.POINTER(TreeLeaf), _i64, C.POINTER(_i64)]),
    "hm_matvec": (_i32, [_vp, _dp, _i64, _dp, _i64, _i32]),
    "hm_matvec_device": (_i32, [_vp, _vp, _vp, _i32, _vp]),
    "hm_matvec_device_allgather": (_i32, [_vp, _vp, C.POINTER(C.c_uint64), _i32, _i32, _i32, _vp]),
    "hm_dist_get_id": (_i32, [_vp]),
    "hm_dist_init": (_i32, [_vp, _vp, _i32, _i32]),
    "hm_dist_buffers": (_i32, [_vp, C.POINTER(_vp), C.POINTER(_vp)]),
    "hm_dist_bcast_x": (_i32, [_vp, _vp, _i32, _i32, _vp]),
    "hm_dist_push_x": (_i32, [_vp, _vp, _i32, _i32, _vp]),
    "hm_dist_matvec_device": (_i32, [_vp, _vp, _i32, _i32, _vp]),
    "hm_dist_barrier": (_i32, [_vp, _vp]),
    "hm_dist_check": (_i32, [_vp]),
    "hm_dist_matvec": (_i32, [_vp, _dp, _i64, _dp, _i64, _i32, _i32]),
    "hm_matvec_adjoint": (_i32, [_vp, _dp, _i64, _dp, _i64, _i32]),
    "hm_matvec_adjoint_device": (_i32, [_vp, _vp, _vp, _i32, _vp]),
    "hm_matmat": (_i32, [_vp, _dp, _i64, _dp, _i64, _i64, _i32]),
    "hm_matmat_device": (_i32, [_vp, _vp, _i64, _vp, _i64, _i64, _i32, _vp]),
    "hm_plan_scale": (_i32, [_vp, _dp, _i64, _i32]),
    "hm_plan_timing_begin": (_i32, [_vp, _i32]),
    "hm_plan_timing_end": (_i32, [_vp, _dp, C.POINTER(_i64)]),
    "hm_plan_launches_per_matvec": (_i32, [_vp]),
    "hm_debug_fail_alloc": (_i32, [_i64]),
    "hm_plan_num_leaves": (_i32, [_vp, C.POINTER(_i64)]),
    "hm_plan_leaf_info": (_i32, [_vp, _i64, C.POINTER(_i32), C.POINTER(_i64), C.POINTER(_i64),
                                 C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "hm_plan_read_leaf": (_i32, [_vp, _i64, _i32, _dp, _i64]),
}


def build(verbose: bool = False) -> str:
    """Compile csrc/ for sm_100a with nvcc (csrc/Makefile) into lib/libhmb200.so."""
    cmd = ["make", "-C", CSRC] + ([] if verbose else ["-s"])
    subprocess.check_call(cmd)
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("build did not produce " + LIB_PATH)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`). "
                "There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int):
    if status != HM_OK:
        raise HmError(status, lib().hm_last_error().decode("utf-8", "replace"))
