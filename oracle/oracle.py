"""ctypes binding of the CPU oracle (oracle/srukf_oracle.c).

TEST INFRASTRUCTURE ONLY -- pinned to the reference's own text by tests/test_ref_pin.py (see srukf_oracle.h).  Only tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module; the product package
(cv_monoslam_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsrukf_oracle.so")


class OracleParams(C.Structure):
    _fields_ = [
        ("cam_dx", C.c_double), ("cam_dy", C.c_double), ("cam_cx", C.c_double), ("cam_cy", C.c_double),
        ("cam_k1", C.c_double), ("cam_k2", C.c_double), ("cam_f", C.c_double),
        ("image_width", C.c_int), ("image_height", C.c_int),
        ("a1", C.c_double), ("a2", C.c_double), ("a3", C.c_double), ("a4", C.c_double),
        ("sigma_measure", C.c_double),
        ("weight_type", C.c_int),
        ("alpha", C.c_double), ("beta", C.c_double),
        ("epsilon", C.c_double),
        ("newton_iters", C.c_int),
        ("downdate_mode", C.c_int),
    ]


class OracleWeights(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("gamma", "wm0", "wm0_sr", "wc0", "wc0_sr", "wi", "wi_sr")]


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (gcc -O2)."""
    src = os.path.join(_HERE, "srukf_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return _LIB_PATH


_lib = None
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_bp = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_default_params.argtypes = [C.POINTER(OracleParams)]
        L.oracle_sample_parameters.argtypes = [C.c_int, C.POINTER(OracleParams), C.POINTER(OracleWeights)]
        L.oracle_qr_R.argtypes = [_dp, C.c_int, C.c_int, _dp]
        L.oracle_mchol.argtypes = [_dp, C.c_int, C.c_double, _dp, _dp]
        L.oracle_mchol.restype = C.c_int
        L.oracle_distort.argtypes = [C.POINTER(OracleParams), C.c_double, C.c_double,
                                     C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.oracle_undistort.argtypes = L.oracle_distort.argtypes
        L.oracle_project.argtypes = [C.POINTER(OracleParams), _dp, _dp, C.c_double, _dp,
                                     C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.oracle_odometry_to_control.argtypes = [_dp, _dp, _dp]
        L.oracle_filter_create.argtypes = [C.c_int, C.POINTER(OracleParams)]
        L.oracle_filter_create.restype = C.c_void_p
        L.oracle_filter_destroy.argtypes = [C.c_void_p]
        L.oracle_filter_set_state.argtypes = [C.c_void_p, _dp, _dp]
        L.oracle_filter_get_state.argtypes = [C.c_void_p, _dp, _dp]
        L.oracle_predict_motion.argtypes = [C.c_void_p, _dp]
        L.oracle_predict_measurement.argtypes = [C.c_void_p]
        L.oracle_kalman_update.argtypes = [C.c_void_p, _dp, _bp]
        L.oracle_step.argtypes = [C.c_void_p, _dp, _dp, _bp]
        L.oracle_filter_set_new_features.argtypes = [C.c_void_p, C.c_int]
        L.oracle_filter_get_prediction.argtypes = [C.c_void_p, _dp, _dp, _bp]
        L.oracle_chi2_gate.argtypes = [C.c_void_p, _dp, C.c_double, _bp, _dp]
        L.oracle_init_features.argtypes = [C.POINTER(OracleParams), _dp, _dp, C.c_int, _dp, C.c_double,
                                           C.c_double, _dp, _dp]
        L.oracle_add_features.argtypes = [C.POINTER(OracleParams), C.c_int, _dp, _dp, C.c_int, _dp, C.c_double,
                                          C.c_double, _dp, _dp]
        L.oracle_delete_feature.argtypes = [C.POINTER(OracleParams), C.c_int, _dp, _dp, C.c_int, _dp, _dp]
        L.oracle_batch_step.argtypes = [C.c_int, C.c_int, C.POINTER(OracleParams), _dp, _dp, _dp, _dp, _bp,
                                        C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def default_params(**over) -> OracleParams:
    p = OracleParams()
    lib().oracle_default_params(C.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    return p


def sample_parameters(Na: int, p: OracleParams | None = None) -> dict:
    p = p or default_params()
    w = OracleWeights()
    lib().oracle_sample_parameters(Na, C.byref(p), C.byref(w))
    return {k: getattr(w, k) for k, _ in OracleWeights._fields_}


def qr_R(A: np.ndarray) -> np.ndarray:
    A = np.ascontiguousarray(A, dtype=np.float64)
    m, n = A.shape
    R = np.zeros((n, n))
    lib().oracle_qr_R(A, m, n, R)
    return R


def mchol(G: np.ndarray, epsilon: float = 1e-13):
    G = np.ascontiguousarray(G, dtype=np.float64)
    n = G.shape[0]
    S = np.zeros((n, n))
    E = np.zeros(n)
    nmod = lib().oracle_mchol(G, n, epsilon, S, E)
    return S, E, nmod


def distort(p, ux, uy):
    ox, oy = C.c_double(), C.c_double()
    lib().oracle_distort(C.byref(p), ux, uy, C.byref(ox), C.byref(oy))
    return ox.value, oy.value


def undistort(p, dx, dy):
    ox, oy = C.c_double(), C.c_double()
    lib().oracle_undistort(C.byref(p), dx, dy, C.byref(ox), C.byref(oy))
    return ox.value, oy.value


def project(p, feat6, pos3, theta, err2=(0.0, 0.0)):
    ox, oy = C.c_double(), C.c_double()
    lib().oracle_project(C.byref(p), np.ascontiguousarray(feat6, dtype=np.float64),
                         np.ascontiguousarray(pos3, dtype=np.float64), float(theta),
                         np.ascontiguousarray(err2, dtype=np.float64), C.byref(ox), C.byref(oy))
    return ox.value, oy.value


def odometry_to_control(o0, o1) -> np.ndarray:
    u = np.zeros(3)
    lib().oracle_odometry_to_control(np.ascontiguousarray(o0, dtype=np.float64),
                                     np.ascontiguousarray(o1, dtype=np.float64), u)
    return u


def init_features(p, x4, S4, kp, rho0, sigma_rho):
    kp = np.ascontiguousarray(kp, dtype=np.float64).reshape(-1, 2)
    M = kp.shape[0]
    n = 6 * M + 4
    x = np.zeros(n)
    S = np.zeros((n, n))
    lib().oracle_init_features(C.byref(p), np.ascontiguousarray(x4, dtype=np.float64),
                               np.ascontiguousarray(S4, dtype=np.float64), M, kp, rho0, sigma_rho, x, S)
    return x, S


def add_features(p, x, S, kp, rho0, sigma_rho):
    """integrateFeaturesInformation on a non-empty map: (x [n], S [n,n]) + M key-points -> (x [n+6M], S)"""
    kp = np.ascontiguousarray(kp, dtype=np.float64).reshape(-1, 2)
    M = kp.shape[0]
    n = x.shape[0]
    xo = np.zeros(n + 6 * M)
    So = np.zeros((n + 6 * M, n + 6 * M))
    lib().oracle_add_features(C.byref(p), (n - 4) // 6, np.ascontiguousarray(x, dtype=np.float64),
                              np.ascontiguousarray(S, dtype=np.float64), M, kp, rho0, sigma_rho, xo, So)
    return xo, So


def delete_feature(p, x, S, id_):
    """deleteOneFeature (SLAM.cpp:2637-2663): (x [n], S [n,n]) -> (x [n-6], S [n-6,n-6])"""
    n = x.shape[0]
    L = (n - 4) // 6
    xo = np.zeros(n - 6)
    So = np.zeros((n - 6, n - 6))
    lib().oracle_delete_feature(C.byref(p), L, np.ascontiguousarray(x, dtype=np.float64),
                                np.ascontiguousarray(S, dtype=np.float64), int(id_), xo, So)
    return xo, So


class Filter:
    """One reference filter (materialised sigma matrices, literal update)."""

    def __init__(self, L: int, p: OracleParams | None = None):
        self.p = p or default_params()
        self.L, self.n = L, 6 * L + 4
        self._h = lib().oracle_filter_create(L, C.byref(self.p))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_filter_destroy(self._h)
            self._h = None

    def set_state(self, x, S):
        lib().oracle_filter_set_state(self._h, np.ascontiguousarray(x, dtype=np.float64),
                                      np.ascontiguousarray(S, dtype=np.float64))

    def get_state(self):
        x = np.zeros(self.n)
        S = np.zeros((self.n, self.n))
        lib().oracle_filter_get_state(self._h, x, S)
        return x, S

    def predict_motion(self, u):
        lib().oracle_predict_motion(self._h, np.ascontiguousarray(u, dtype=np.float64))

    def predict_measurement(self):
        lib().oracle_predict_measurement(self._h)

    def kalman_update(self, z, matched):
        lib().oracle_kalman_update(self._h, np.ascontiguousarray(z, dtype=np.float64),
                                   np.ascontiguousarray(matched, dtype=np.uint8))

    def set_new_features(self, n_new: int):
        """m_nAddings: KalmanUpdate takes the NEED_REORDER branch (SLAM.cpp:2083-2086) while it is non-zero"""
        lib().oracle_filter_set_new_features(self._h, int(n_new))

    def prediction(self):
        """(m_allPredictSet [L,2], Si [L,2,2], isVisible [L]) after predict_measurement"""
        hbar = np.zeros((self.L, 2))
        si = np.zeros((self.L, 2, 2))
        vis = np.zeros(self.L, dtype=np.uint8)
        lib().oracle_filter_get_prediction(self._h, hbar, si, vis)
        return hbar, si, vis

    def chi2_gate(self, z, threshold=5.99146454710798):
        acc = np.zeros(self.L, dtype=np.uint8)
        d2 = np.zeros(self.L)
        lib().oracle_chi2_gate(self._h, np.ascontiguousarray(z, dtype=np.float64), float(threshold), acc, d2)
        return acc, d2

    def step(self, u, z, matched):
        lib().oracle_step(self._h, np.ascontiguousarray(u, dtype=np.float64),
                          np.ascontiguousarray(z, dtype=np.float64),
                          np.ascontiguousarray(matched, dtype=np.uint8))


def batch_step(p, x, S, u, z, matched, nthreads=0):
    """Advance B filters by nsteps.  x [B,n], S [B,n,n] are updated in place.
    u [nsteps,B,3], z [nsteps,B,L,2], matched [nsteps,B,L].  Returns max |E| per filter."""
    B, n = x.shape
    L = (n - 4) // 6
    nsteps = u.shape[0]
    assert x.flags.c_contiguous and S.flags.c_contiguous
    u = np.ascontiguousarray(u, dtype=np.float64)
    z = np.ascontiguousarray(z, dtype=np.float64)
    matched = np.ascontiguousarray(matched, dtype=np.uint8)
    maxE = np.zeros(B)
    lib().oracle_batch_step(B, L, C.byref(p), x, S, u, z, matched, 0, nsteps, nthreads,
                            maxE.ctypes.data_as(C.c_void_p))
    return maxE
