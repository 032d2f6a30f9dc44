"""ctypes binding of include/srukf.h (libsrukf_b200.so).  Plumbing only -- no arithmetic here."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.environ.get("SRUKF_LIB_PATH") or os.path.join(_HERE, "libsrukf_b200.so")   # override: tuning builds only

SRUKF_OK, SRUKF_EINVAL, SRUKF_ECUDA, SRUKF_ENOMEM, SRUKF_ESTATE, SRUKF_ENODEV = 0, -1, -2, -3, -4, -5
FLAG_NAN, FLAG_GMW_FLOOR, FLAG_GMW_MODIFIED, FLAG_OUT_OF_VIEW, FLAG_INVISIBLE, FLAG_FALLBACK = 1, 2, 4, 8, 16, 32


class SrukfError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"srukf error {code}: {msg}")
        self.code = code


class SrukfParams(C.Structure):
    """Mirror of `SrukfParams` in include/srukf.h (CSLAM scalar members, SLAM.cpp:158-343)."""
    _fields_ = [
        ("cam_dx", C.c_double), ("cam_dy", C.c_double), ("cam_cx", C.c_double), ("cam_cy", C.c_double),
        ("cam_k1", C.c_double), ("cam_k2", C.c_double), ("cam_f", C.c_double),
        ("image_width", C.c_int32), ("image_height", C.c_int32),
        ("a1", C.c_double), ("a2", C.c_double), ("a3", C.c_double), ("a4", C.c_double),
        ("sigma_measure", C.c_double),
        ("weight_type", C.c_int32),
        ("alpha", C.c_double), ("beta", C.c_double),
        ("epsilon", C.c_double),
        ("newton_iters", C.c_int32),
        ("downdate_mode", C.c_int32),
    ]


# every symbol include/srukf.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
SYMBOLS = {
    "srukf_default_params": (None, [C.POINTER(SrukfParams)]),
    "srukf_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(SrukfParams), C.POINTER(_VP)]),
    "srukf_destroy": (C.c_int, [_VP]),
    "srukf_set_state": (C.c_int, [_VP, _VP, _VP]),
    "srukf_get_state": (C.c_int, [_VP, _VP, _VP]),
    "srukf_set_state_dense": (C.c_int, [_VP, _VP, _VP]),
    "srukf_get_state_dense": (C.c_int, [_VP, _VP, _VP]),
    "srukf_predict_motion": (C.c_int, [_VP, _VP]),
    "srukf_predict_measurement": (C.c_int, [_VP]),
    "srukf_get_prediction": (C.c_int, [_VP, _VP, _VP, _VP]),
    "srukf_init_features": (C.c_int, [_VP, _VP, _VP, _VP, C.c_double, C.c_double]),
    "srukf_kalman_update_reorder": (C.c_int, [_VP, _VP, _VP, C.c_int]),
    "srukf_add_features": (C.c_int, [_VP, _VP, _VP, C.c_double, C.c_double]),
    "srukf_delete_feature": (C.c_int, [_VP, _VP, _VP]),
    "srukf_chi2_gate": (C.c_int, [_VP, _VP, C.c_double, _VP, _VP]),
    "srukf_kalman_update": (C.c_int, [_VP, _VP, _VP]),
    "srukf_step": (C.c_int, [_VP, _VP, _VP, _VP]),
    "srukf_get_x_async": (C.c_int, [_VP, _VP]),
    "srukf_mchol": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_double, _VP, _VP, _VP]),
    "srukf_qr_R": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _VP, _VP]),
    "srukf_generate_sigma_points": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_double, _VP, _VP, _VP]),
    "srukf_cholesky_update": (C.c_int, [_VP, _VP, C.c_int, C.c_int, C.c_int, C.c_int]),
    "srukf_fp64_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "srukf_step_dev": (C.c_int, [_VP, _VP, _VP, _VP]),
    "srukf_set_state_dev": (C.c_int, [_VP, C.c_int, C.c_int, _VP, _VP]),
    "srukf_get_cov_block": (C.c_int, [_VP, C.c_int, C.c_int, _VP]),
    "srukf_get_flags": (C.c_int, [_VP, _VP]),
    "srukf_clear_flags": (C.c_int, [_VP]),
    "srukf_stats": (C.c_int, [_VP, _VP, _VP]),
    "srukf_sync": (C.c_int, [_VP]),
    "srukf_stream": (C.c_int, [_VP, C.POINTER(C.c_uint64)]),
    "srukf_launch_count": (C.c_int, [_VP, C.POINTER(C.c_uint64)]),
    "srukf_set_profiling": (C.c_int, [_VP, C.c_int]),
    "srukf_get_kernel_times": (C.c_int, [_VP, _VP, _VP]),
    "srukf_get_phase_cycles": (C.c_int, [_VP, _VP]),
    "srukf_last_error": (C.c_char_p, []),
    "srukf_version": (C.c_char_p, []),
}

_lib = None


def lib_path() -> str:
    return _LIB


def load_library():
    """Load libsrukf_b200.so; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            raise SrukfError(SRUKF_ENODEV, f"{_LIB} is missing: run `python -c 'import __graft_entry__ as g; "
                                           "g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(_LIB)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int):
    if rc != SRUKF_OK:
        raise SrukfError(rc, load_library().srukf_last_error().decode())


def default_params(**over) -> SrukfParams:
    p = SrukfParams()
    load_library().srukf_default_params(C.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    return p


def ptr(a) -> int | None:
    """Host numpy array (C-contiguous) or integer device address -> void*."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return int(a)
    assert isinstance(a, np.ndarray) and a.flags.c_contiguous, "need a C-contiguous numpy array"
    return a.ctypes.data
