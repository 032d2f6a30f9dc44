"""ctypes binding of oracle/_ref/libsrukf_ref.so: the reference's OWN SRUKF function bodies (extracted verbatim from
/root/reference/MonoSLAM/SLAM.cpp at build time, compiled against oracle/ref_shim/) behind oracle/ref_shim/ref_driver.cpp.

TEST INFRASTRUCTURE ONLY.  It pins oracle/srukf_oracle.c to the reference's text (tests/test_ref_pin.py) and serves as
the `"kind": "reference"` CPU baseline of bench.py.  The product package never imports it.

The library is built where /root/reference exists (this container) and travels to the GPU box as a prebuilt,
git-ignored file; `available()` is False where neither exists.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libsrukf_ref.so")
REF_DIR = os.environ.get("SRUKF_REFERENCE_DIR", "/root/reference/MonoSLAM")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_bp = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_lib = None

UPDATING, DOWNDATING = 0, 1            # FLAG_4_UPDATING / FLAG_4_DOWNDATING, SLAM.cpp:31-32
NEED_REORDER, NEEDNOT_REORDER = 0, 1   # SLAM.cpp:36-37


def build(force: bool = False) -> str | None:
    """Extract + compile when the reference tree is present; otherwise keep whatever prebuilt library exists."""
    if os.path.exists(os.path.join(REF_DIR, "SLAM.cpp")):
        srcs = [os.path.join(REF_DIR, "SLAM.cpp"), os.path.join(_HERE, "ref_shim", "ref_shim.h"),
                os.path.join(_HERE, "ref_shim", "ref_driver.cpp"), os.path.join(_HERE, "ref_shim", "extract_ref.py"),
                os.path.join(_HERE, "srukf_oracle.c")]
        stale = not os.path.exists(_LIB_PATH) or any(os.path.getmtime(_LIB_PATH) < os.path.getmtime(s) for s in srcs)
        if force or stale:
            os.makedirs(os.path.join(_HERE, "_ref"), exist_ok=True)
            subprocess.check_call(["make", "-s", "-C", _HERE, "_ref", f"REF={REF_DIR}"])
    return _LIB_PATH if os.path.exists(_LIB_PATH) else None


def available() -> bool:
    try:
        return build() is not None
    except Exception:
        return os.path.exists(_LIB_PATH)


def manifest() -> list[tuple[str, int, int, str]]:
    """(function, first line, last line, sha1[:12]) of every extracted body"""
    out = []
    with open(os.path.join(_HERE, "_ref", "manifest.txt")) as f:
        for line in f:
            name, a, b, h = line.split()
            out.append((name, int(a), int(b), h))
    return out


def lib():
    global _lib
    if _lib is None:
        if build() is None:
            raise RuntimeError("oracle/_ref/libsrukf_ref.so is absent and the reference tree is not here to build it")
        L = C.CDLL(_LIB_PATH)
        vp, ci, cd = C.c_void_p, C.c_int, C.c_double
        L.ref_create.restype = vp
        L.ref_destroy.argtypes = [vp]
        L.ref_get_params.argtypes = [vp, _dp]
        L.ref_set_weight_type.argtypes = [vp, ci]
        L.ref_set_camera.argtypes = [vp, cd, cd]
        L.ref_sample_parameters.argtypes = [vp, ci, _dp]
        L.ref_mchol.argtypes = [vp, _dp, ci, _dp]
        L.ref_qr_R.argtypes = [vp, _dp, ci, ci, _dp]
        L.ref_distort.argtypes = [vp, cd, cd, C.POINTER(cd), C.POINTER(cd)]
        L.ref_undistort.argtypes = [vp, cd, cd, C.POINTER(cd), C.POINTER(cd)]
        L.ref_project.argtypes = [vp, _dp, _dp, cd, _dp, _dp]
        L.ref_set_state.argtypes = [vp, ci, _dp, _dp]
        L.ref_state_dim.argtypes = [vp]
        L.ref_state_dim.restype = ci
        L.ref_get_state.argtypes = [vp, _dp, _dp]
        L.ref_predict_motion.argtypes = [vp, _dp, _dp, _dp, _dp]
        L.ref_get_sigma.argtypes = [vp, vp, C.POINTER(ci), C.POINTER(ci)]
        L.ref_predict_measurement.argtypes = [vp]
        L.ref_get_prediction.argtypes = [vp, _dp, _dp, _bp, vp]
        L.ref_kalman_update.argtypes = [vp, _dp, _bp, ci]
        L.ref_cholesky_update.argtypes = [vp, ci, _dp, _dp, ci, ci, ci, ci, ci, _dp]
        L.ref_get_permutation.argtypes = [vp, _dp]
        L.ref_delete_feature.argtypes = [vp, ci]
        L.ref_init_features.argtypes = [vp, ci, _dp, cd, cd]
        _lib = L
    return _lib


PARAM_NAMES = ("cam_dx cam_dy cam_cx cam_cy cam_k1 cam_k2 cam_f cam_f1 cam_f2 a1 a2 a3 a4 sigma_measure rho sigma_rho "
               "sigma_x sigma_y sigma_z sigma_theta epsilon image_width image_height weight_type noise_type alpha beta").split()


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def control_to_odometry(u):
    """Two odometry poses (x, y, theta) whose SLAM.cpp:1446-1450 control is u up to rounding; the exact control the
    reference derives comes back from Slam.predict_motion."""
    r1, t, r2 = (float(v) for v in u)
    return np.zeros(3), np.array([t * np.cos(r1), t * np.sin(r1), r1 + r2])


class Slam:
    """One CSLAM object of the reference (constructed by its own constructor + initializeParameters)."""

    def __init__(self, weight_type: int | None = None):
        self._h = lib().ref_create()
        if weight_type is not None:
            lib().ref_set_weight_type(self._h, int(weight_type))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ref_destroy(self._h)
            self._h = None

    def params(self) -> dict:
        v = np.zeros(27)
        lib().ref_get_params(self._h, v)
        return dict(zip(PARAM_NAMES, v.tolist()))

    def set_camera(self, k1, k2):
        lib().ref_set_camera(self._h, float(k1), float(k2))

    def sample_parameters(self, Na: int) -> dict:
        v = np.zeros(7)
        lib().ref_sample_parameters(self._h, int(Na), v)
        return dict(zip(("gamma", "wm0", "wm0_sr", "wc0", "wc0_sr", "wi", "wi_sr"), v.tolist()))

    def mchol(self, G):
        G = _c(G)
        S = np.zeros_like(G)
        lib().ref_mchol(self._h, G, G.shape[0], S)
        return S

    def qr_R(self, A):
        A = _c(A)
        R = np.zeros((A.shape[1], A.shape[1]))
        lib().ref_qr_R(self._h, A, A.shape[0], A.shape[1], R)
        return R

    def distort(self, ux, uy):
        a, b = C.c_double(), C.c_double()
        lib().ref_distort(self._h, ux, uy, C.byref(a), C.byref(b))
        return a.value, b.value

    def undistort(self, dx, dy):
        a, b = C.c_double(), C.c_double()
        lib().ref_undistort(self._h, dx, dy, C.byref(a), C.byref(b))
        return a.value, b.value

    def project(self, feat6, pos3, theta, err2=(0.0, 0.0)):
        out = np.zeros(2)
        lib().ref_project(self._h, _c(feat6), _c(pos3), float(theta), _c(err2), out)
        return out[0], out[1]

    def set_state(self, x, S):
        x = _c(x)
        L = (x.shape[0] - 4) // 6
        lib().ref_set_state(self._h, L, x, _c(S))

    @property
    def n(self):
        return lib().ref_state_dim(self._h)

    def get_state(self):
        n = self.n
        x, S = np.zeros(n), np.zeros((n, n))
        lib().ref_get_state(self._h, x, S)
        return x, S

    def predict_motion_odometry(self, odo_prev, odo_now):
        """predictMotion from two odometry poses; returns (Ut, diag Mt) as the reference derived them"""
        u, m = np.zeros(3), np.zeros(3)
        lib().ref_predict_motion(self._h, _c(odo_prev), _c(odo_now), u, m)
        return u, m

    def sigma(self):
        na, p = C.c_int(), C.c_int()
        lib().ref_get_sigma(self._h, None, C.byref(na), C.byref(p))
        s = np.zeros((na.value, p.value))
        lib().ref_get_sigma(self._h, s.ctypes.data_as(C.c_void_p), None, None)
        return s

    def predict_measurement(self):
        lib().ref_predict_measurement(self._h)

    def prediction(self):
        L = (self.n - 4) // 6
        hbar, si, vis = np.zeros((L, 2)), np.zeros((L, 2, 2)), np.zeros(L, dtype=np.uint8)
        lib().ref_get_prediction(self._h, hbar, si, vis, None)
        return hbar, si, vis

    def kalman_update(self, z, matched, n_new: int = 0):
        lib().ref_kalman_update(self._h, _c(z), np.ascontiguousarray(matched, dtype=np.uint8), int(n_new))

    def cholesky_update(self, S, U, mode, order, n_new=0, nmap=0):
        S, U = _c(S), _c(U)
        n = S.shape[0]
        U = U.reshape(n, -1)
        out = np.zeros_like(S)
        lib().ref_cholesky_update(self._h, n, S, U, U.shape[1], int(mode), int(order), int(n_new), int(nmap), out)
        return out

    def delete_feature(self, id_):
        lib().ref_delete_feature(self._h, int(id_))

    def init_features(self, kp, rho, sigma_rho):
        kp = _c(kp).reshape(-1, 2)
        lib().ref_init_features(self._h, kp.shape[0], kp, float(rho), float(sigma_rho))
