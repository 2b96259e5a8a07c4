"""ctypes binding of libuglad_b200.so (the C-ABI declared in include/uglad_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, an exception is
raised.  Build it with `python -c "import __graft_entry__ as g; g.build()"` or `make`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# UGLAD_B200_LIB: developer override (A/B of two builds of the library in one GPU session)
LIB_PATH = os.environ.get("UGLAD_B200_LIB") or os.path.join(_HERE, "lib", "libuglad_b200.so")


class UgladDims(C.Structure):
    _fields_ = [("B", C.c_int), ("D", C.c_int), ("L", C.c_int), ("H", C.c_int),
                ("init_diag", C.c_int), ("B_total", C.c_int), ("exact_sqrt", C.c_int),
                ("lambda_init", C.c_float)]


class UgladPeers(C.Structure):
    _fields_ = [("world", C.c_int), ("rank", C.c_int), ("tag", C.c_uint), ("slots", C.c_void_p * 8),
                ("tag_dev", C.c_void_p)]


# name -> (restype, argtypes); every symbol include/uglad_b200.h declares
_P, _I, _F, _Z = C.c_void_p, C.c_int, C.c_float, C.c_size_t
_DP = C.POINTER(UgladDims)
SIGNATURES = {
    "uglad_abi_version": (_I, []),
    "uglad_last_error": (C.c_char_p, []),
    "uglad_param_count": (_Z, [_I]),
    "uglad_covariance": (_I, [_P, _I, _I, _I, _P, _P, _P]),
    "uglad_covariance_scratch_floats": (_Z, [_I, _I, _I]),
    "uglad_covariance_ws": (_I, [_P, _I, _I, _I, _P, _P, _P, _P]),
    "uglad_eigh_scratch_floats": (_Z, [_I, _I]),
    "uglad_eigh": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _P]),
    "uglad_condition_covariance": (_I, [_P, _I, _I, _F, _P, _P, _P, _P, _P]),
    "uglad_eigh_warm": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "uglad_condition_covariance_warm": (_I, [_P, _I, _I, _F, _P, _P, _P, _P, _P, _P, _P]),
    "uglad_condition_scratch_floats": (_Z, [_I, _I]),
    "uglad_condition_covariance_x": (_I, [_P, _P, _P, _I, _I, _I, _F, _P, _P, _P, _P, _P, _P, _P]),
    "uglad_small_d_max": (_I, []),
    "uglad_eig_path": (_I, [_I, _I]),
    "uglad_workspace_floats": (_Z, [_DP]),
    "uglad_workspace_offset": (_Z, [_DP, C.c_char_p]),
    "uglad_glad_init_forward": (_I, [_DP, _P, _P, _P, _P, _P, _P]),
    "uglad_glad_layer_forward": (_I, [_DP, _I, _P, _P, _P, _P, _P]),
    "uglad_glad_forward": (_I, [_DP, _P, _P, _P, _P, _P, _P, _P]),
    "uglad_glad_backward": (_I, [_DP, _P, _P, _P, _P, _P, _P, _P, _P]),
    "uglad_peer_slots_bytes": (_Z, [_I]),
    "uglad_glad_forward_sharded": (_I, [_DP, _P, _P, _P, _P, _P, _P, C.POINTER(UgladPeers), _P]),
    "uglad_peer_alloc": (_I, [_Z, C.POINTER(C.c_void_p), C.c_char_p]),
    "uglad_peer_open": (_I, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "uglad_peer_close": (_I, [_P, _I]),
    "uglad_loss_scratch_floats": (_Z, [_I, _I]),
    "uglad_glasso_loss": (_I, [_P, _P, _I, _I, _I, _F, _P, _P, _P, _P]),
    "uglad_glasso_loss_prior": (_I, [_P, _P, _P, _I, _I, _I, _F, _P, _P, _P, _P]),
    "uglad_launch_count": (C.c_ulonglong, []),
    "uglad_profile": (_I, [_I, C.POINTER(C.c_double), C.POINTER(C.c_ulonglong)]),
    "uglad_profile_read": (_I, [_I, C.POINTER(C.c_double), C.POINTER(C.c_ulonglong), C.POINTER(C.c_double)]),
    "uglad_tune": (_I, [C.c_char_p, _I]),
    "uglad_tc_debug_buffer": (_I, [_P]),
    "uglad_tc_gemm_repeat": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "uglad_tc_gemm_scratch_floats": (_Z, [_I, _I, _I, _I]),
    "uglad_tc_gemm": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _F, _P, _P]),
    "uglad_z_update": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
}

_lib = None


class UgladError(RuntimeError):
    pass


def load():
    """Load the shared library once; raise if it is absent (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise UgladError(
            f"{LIB_PATH} not found: build the CUDA library first (make, or "
            "__graft_entry__.build()).  uglad_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype, fn.argtypes = res, args
    if lib.uglad_abi_version() != 1:
        raise UgladError("libuglad_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        raise UgladError(f"{what}: {load().uglad_last_error().decode()}")
