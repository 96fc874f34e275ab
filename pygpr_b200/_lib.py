"""ctypes binding of libpygpr_b200.so (the C ABI in include/pygpr_b200.h).

There is NO fallback: if the shared library is missing or a call fails, this raises.  The library is
built in-tree by ``pygpr_b200/csrc/build.py`` (``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_int, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpygpr_b200.so")

TILE = 128
ABI_VERSION = 8

# name -> (restype, argtypes).  Pointers are passed as integers (tensor.data_ptr()).
_P, _I, _LL = c_void_p, c_int, c_longlong
SIGNATURES = {
    "gpr_abi_version": (c_int, []),
    "gpr_last_error": (c_char_p, []),
    "gpr_set_option": (c_int, [c_char_p, c_char_p]),
    "gpr_device_sm_count": (c_int, []),
    "gpr_predict_scratch_bytes": (c_longlong, [_I]),
    "gpr_tril_inverse": (c_int, [_P, _P, _I, _I, _P, _P, _P]),
    "gpr_predict_gemm": (c_int, [_P, _P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _I, _P, _P]),
    "gpr_factor_partials": (c_int, [_P, _I, _P, _I, _P, _P]),
    "gpr_predict_gemm_scratch_bytes": (c_longlong, [_I, _I, _I, _I]),
    "gpr_predict_keep": (c_int, [_P, _P, _P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "gpr_predict_shared": (c_int, [_P, _P, _P, _P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P, _I, _P, _P]),
    "gpr_lml_grad_scratch_bytes": (c_longlong, [_I, _I, _I]),
    "gpr_lml_grad_shared": (c_int, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _I, _P, _I]),
    "gpr_se_scale": (c_int, [_P, _LL, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    "gpr_se_build": (c_int, [_P, _P, _P, _I, _P, _I, _I, _I, _I, _P, _I, _I, _I, _P]),
    "gpr_se_cross": (c_int, [_P, _P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _I, _P, _P]),
    "gpr_se_cross_ld": (c_int, [_P, _P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _I, _P, _LL, _LL, _P]),
    "gpr_se_cross_planes_ok": (c_int, [_I, _I]),
    "gpr_comm_unique_id_bytes": (c_int, []),
    "gpr_comm_unique_id": (c_int, [_P]),
    "gpr_comm_init": (c_int, [_P, _I, _I, _P]),
    "gpr_comm_allreduce_f64": (c_int, [_P, _P, _LL, _P]),
    "gpr_comm_destroy": (c_int, [_P]),
    "gpr_se_cross_planes": (c_int, [_P, _P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _I, c_double, _P, _P, _P]),
    "gpr_se_kernel_grad": (c_int, [_P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "gpr_potrf_batched": (c_int, [_P, _I, _P, _P, _P, _I, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gpr_potrf_work_bytes": (c_longlong, [_I, _I]),
    "gpr_potrf_grbcm": (c_int, [_P, _P, _P, _P, _I, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P,
                                _P, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gpr_potrf_grbcm_work_bytes": (c_longlong, [_I, _I, _I]),
    "gpr_copy_leading_block": (c_int, [_P, _I, _I, _P, _I, _I, _I, _P]),
    "gpr_potrf_batched_shared": (c_int, [_P, _P, _P, _P, _I, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _I, _P]),
    "gpr_trsv_backward": (c_int, [_P, _P, _I, _I, _P, _P, _P, _P]),
    "gpr_trsv_work_bytes": (c_longlong, [_I]),
    "gpr_predict": (c_int, [_P, _P, _P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P]),
    "gpr_lml_grad": (c_int, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    "gpr_predict_full_scratch_bytes": (c_longlong, [_I, _I, _I]),
    "gpr_potri_scratch_bytes": (c_longlong, [_I, _I]),
    "gpr_predict_full": (c_int, [_P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "gpr_potri_batched": (c_int, [_P, _P, _I, _I, _I, _P, _I, _P, _P]),
    "grbcm_full_combine": (c_int, [_P, _P, _I, _I, _P, _P]),
    "gpr_predict_dense": (c_int, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P]),
    "gpr_predict_dense_scratch_bytes": (c_longlong, [_I, _I, _I]),
    "gpr_covar_dense": (c_int, [_P, _P, _I, _I, _I, _P, _P]),
    "gpr_trace_grad": (c_int, [_P, _P, _I, _P, _I, _I, _I, _P, _P, _P]),
    "gpr_trace_grad_scratch_bytes": (c_longlong, [_I, _I, _I]),
    "gpr_sq_dist": (c_int, [_P, _P, _I, _P, _P, _I, _I, _I, _I, _P, _P]),
    "grbcm_partial": (c_int, [_P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "grbcm_finalize": (c_int, [_P, _P, _P, _I, _P, _P, _P, _P, _P]),
    "gpr_trtri_batched": (c_int, [_P, _P, _I, _I, _P, _P, _I, _I]),
    "gpr_factor_window": (c_int, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "gpr_tril_inverse_rows": (c_int, [_P, _P, _I, _I, _I, _P, _P]),
    "gpr_vt_reduce": (c_int, [_P, _LL, _P, _LL, _P, _P, _I, _I, _I, _P, _P, _P]),
    "gpr_ozaki_rows_padded": (c_longlong, [_I]),
    "gpr_ozaki_k_padded": (c_longlong, [_I]),
    "gpr_ozaki_planes_bytes": (c_longlong, [_I, _I, _I]),
    "gpr_ozaki_split": (c_int, [_P, _LL, _I, _I, _I, c_double, _P, _P, _I, _LL, _P]),
    "gpr_ozaki_gemm_nt": (c_int, [_P, _P, _I, _P, _P, _I, _I, _I, _I, c_double, _P, _LL, _I, _LL, _I, _I, c_double, _I, _P]),
    "gpr_ozaki_gemm_nt_window": (c_int, [_P, _P, _I, _I, _I, _P, _P, _I, _I, _I, _I, _I, _I, c_double, _P, _LL, _I, _LL, _I, _I, c_double, _I, _P]),
    "gpr_gemv_sub": (c_int, [_P, _LL, _LL, _I, _I, _I, _P, _LL, _P, _LL, _P]),
    "gpr_triu_rows": (c_int, [_P, _P, _I, _I, _P, _P]),
    "gpr_grad_from_kinv": (c_int, [_P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
}

_lib = None


class GprError(RuntimeError):
    """A C-ABI call returned a negative status."""


def load() -> ctypes.CDLL:
    """Load the shared library (once).  Raises if it has not been built — never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("PYGPR_B200_LIB", LIB_PATH)   # developer knob: A/B a differently built library
    if not os.path.exists(path):
        raise ImportError(
            "pygpr_b200: %s not found. Build it with `python pygpr_b200/csrc/build.py` "
            "(or __graft_entry__.build()). There is no CPU fallback." % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means the .so is stale
        fn.restype = res
        fn.argtypes = args
    if lib.gpr_abi_version() != ABI_VERSION:
        raise ImportError("pygpr_b200: ABI version mismatch, rebuild the library")
    _lib = lib
    return lib


def call(name: str, *args) -> None:
    """Invoke an int-returning entry point; raise GprError on a negative status."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.gpr_last_error()
        raise GprError("%s failed (%d): %s" % (name, rc, msg.decode() if msg else ""))
