"""Host-side mirror of the reference's CSLAM interface for the SRUKF path, batched over B filters.

Method names follow MonoSLAM/SLAM.h:359-360,372 (`predictMotion`, `predictMeasurement`,
`KalmanUpdate`, `SLAM`); the data members the reference exchanges through (`m_X_k`, `m_S_k`,
`m_P_k`, `Ut`, `m_allPredictSet`, map_p->Si / isVisible / matchLocation / isMatching) become
explicit arguments and getters.  All arithmetic happens in libsrukf_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


def tri_pack(S: np.ndarray) -> np.ndarray:
    """[B,n,n] upper-triangular -> packed row-major [B, n(n+1)/2] (layout of include/srukf.h)."""
    n = S.shape[-1]
    iu = np.triu_indices(n)
    return np.ascontiguousarray(S[..., iu[0], iu[1]])


def tri_unpack(Sp: np.ndarray, n: int) -> np.ndarray:
    iu = np.triu_indices(n)
    out = np.zeros(Sp.shape[:-1] + (n, n))
    out[..., iu[0], iu[1]] = Sp
    return out


class CSLAMBatch:
    """B independent CSLAM filters (SLAM.h:118) with L landmarks each on one CUDA device."""

    def __init__(self, B: int, L: int, params: capi.SrukfParams | None = None, device: int = 0):
        self._lib = capi.load_library()
        self.B, self.L, self.n = int(B), int(L), 6 * int(L) + 4
        self.ntri = self.n * (self.n + 1) // 2
        self.params = params or capi.default_params()
        self.device = int(device)
        h = C.c_void_p()
        capi.check(self._lib.srukf_create(device, self.B, self.L, C.byref(self.params), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.srukf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- m_X_k / m_S_k ------------------------------------------------------------------------
    def set_state(self, x: np.ndarray, S: np.ndarray):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(self.B, self.n)
        S = np.ascontiguousarray(S, dtype=np.float64)
        if S.shape == (self.B, self.ntri):
            capi.check(self._lib.srukf_set_state(self._h, capi.ptr(x), capi.ptr(S)))
        else:
            S = S.reshape(self.B, self.n, self.n)
            capi.check(self._lib.srukf_set_state_dense(self._h, capi.ptr(x), capi.ptr(S)))

    def get_state(self, dense: bool = True):
        x = np.empty((self.B, self.n))
        if dense:
            S = np.empty((self.B, self.n, self.n))
            capi.check(self._lib.srukf_get_state_dense(self._h, capi.ptr(x), capi.ptr(S)))
        else:
            S = np.empty((self.B, self.ntri))
            capi.check(self._lib.srukf_get_state(self._h, capi.ptr(x), capi.ptr(S)))
        return x, S

    @property
    def m_X_k(self) -> np.ndarray:
        return self.get_state()[0]

    @property
    def m_S_k(self) -> np.ndarray:
        return self.get_state()[1]

    def m_P_k(self, r0: int = None, nr: int = None) -> np.ndarray:
        """Block of m_P_k = S^T S (SLAM.cpp:2404); default: the 4x4 robot block."""
        if r0 is None:
            r0, nr = self.n - 4, 4
        out = np.empty((self.B, nr, nr))
        capi.check(self._lib.srukf_get_cov_block(self._h, r0, nr, capi.ptr(out)))
        return out

    # ---- the path -----------------------------------------------------------------------------
    def predictMotion(self, Ut: np.ndarray):
        Ut = np.ascontiguousarray(Ut, dtype=np.float64).reshape(self.B, 3)
        capi.check(self._lib.srukf_predict_motion(self._h, capi.ptr(Ut)))

    def predictMeasurement(self):
        capi.check(self._lib.srukf_predict_measurement(self._h))

    def prediction(self):
        """(m_allPredictSet [B,L,2], Si [B,L,2,2], isVisible [B,L])"""
        hbar = np.empty((self.B, self.L, 2))
        si = np.empty((self.B, self.L, 2, 2))
        vis = np.empty((self.B, self.L), dtype=np.uint8)
        capi.check(self._lib.srukf_get_prediction(self._h, capi.ptr(hbar), capi.ptr(si), capi.ptr(vis)))
        return hbar, si, vis

    def initFeatures(self, x4: np.ndarray, S4: np.ndarray, keypoints: np.ndarray, rho0: float = 1.0 / 3.0,
                     sigma_rho: float | None = None):
        """Frame-1 feature initialisation on the device (addFeatures with an empty map, SLAM.cpp:818-871,
        1177-1334): x4 [B,4], S4 [B,4,4], keypoints [B,L,2] distorted pixels.  Replaces the state of every filter."""
        x4 = np.ascontiguousarray(np.broadcast_to(np.asarray(x4, dtype=np.float64), (self.B, 4)))
        S4 = np.ascontiguousarray(np.broadcast_to(np.asarray(S4, dtype=np.float64), (self.B, 4, 4)))
        kp = np.ascontiguousarray(keypoints, dtype=np.float64).reshape(self.B, self.L, 2)
        if sigma_rho is None:
            sigma_rho = rho0 / 2.0     # SLAM.cpp:173
        capi.check(self._lib.srukf_init_features(self._h, capi.ptr(x4), capi.ptr(S4), capi.ptr(kp), float(rho0),
                                                 float(sigma_rho)))

    def addFeatures(self, keypoints: np.ndarray, rho0: float = 1.0 / 3.0, sigma_rho: float | None = None) -> "CSLAMBatch":
        """integrateFeaturesInformation on a non-empty map (SLAM.cpp:818-871): keypoints [B,M,2] are appended to
        every filter.  Returns a new batch with L+M features; call KalmanUpdateReorder(..., M) on its next frame."""
        kp = np.ascontiguousarray(keypoints, dtype=np.float64)
        kp = kp.reshape(self.B, -1, 2)
        M = kp.shape[1]
        if sigma_rho is None:
            sigma_rho = rho0 / 2.0
        out = CSLAMBatch(self.B, self.L + M, self.params, self.device)
        capi.check(self._lib.srukf_add_features(self._h, out._h, capi.ptr(kp), float(rho0), float(sigma_rho)))
        return out

    def deleteFeature(self, ids) -> "CSLAMBatch":
        """deleteOneFeature (SLAM.cpp:2637-2663) for every filter: filter b drops feature ids[b].  Returns a new
        batch with L-1 features (a handle has a fixed state dimension); this one is left untouched."""
        ids = np.ascontiguousarray(np.broadcast_to(np.asarray(ids, dtype=np.int32), (self.B,)))
        out = CSLAMBatch(self.B, self.L - 1, self.params, self.device)
        capi.check(self._lib.srukf_delete_feature(self._h, out._h, capi.ptr(ids)))
        return out

    CHI2INV_95_2 = 5.99146454710798   # CHI2INV_TABLE(0,2), SLAM.cpp:54

    def chi2Gate(self, candidates: np.ndarray, threshold: float = CHI2INV_95_2):
        """Chi-square gate of dataAssociation (SLAM.cpp:1946-1977): (isMatching mask [B,L], Mahalanobis d2 [B,L])."""
        z = np.ascontiguousarray(candidates, dtype=np.float64).reshape(self.B, self.L, 2)
        acc = np.empty((self.B, self.L), dtype=np.uint8)
        d2 = np.empty((self.B, self.L))
        capi.check(self._lib.srukf_chi2_gate(self._h, capi.ptr(z), float(threshold), capi.ptr(acc), capi.ptr(d2)))
        return acc, d2

    def KalmanUpdate(self, matchLocation: np.ndarray, isMatching: np.ndarray):
        z = np.ascontiguousarray(matchLocation, dtype=np.float64).reshape(self.B, self.L, 2)
        m = np.ascontiguousarray(isMatching, dtype=np.uint8).reshape(self.B, self.L)
        capi.check(self._lib.srukf_kalman_update(self._h, capi.ptr(z), capi.ptr(m)))

    def KalmanUpdateReorder(self, matchLocation: np.ndarray, isMatching: np.ndarray, n_new: int):
        """KalmanUpdate on the frame after n_new features were added (m_nAddings != 0): NEED_REORDER branch of
        GSLCholeskyUpdate (SLAM.cpp:2083-2086, 2122-2138, 2158-2179)."""
        z = np.ascontiguousarray(matchLocation, dtype=np.float64).reshape(self.B, self.L, 2)
        m = np.ascontiguousarray(isMatching, dtype=np.uint8).reshape(self.B, self.L)
        capi.check(self._lib.srukf_kalman_update_reorder(self._h, capi.ptr(z), capi.ptr(m), int(n_new)))

    def SLAM(self, Ut, matchLocation, isMatching):
        """One frame of the hot path: predictMotion + predictMeasurement + KalmanUpdate (SLAM.cpp:91-99)."""
        u = np.ascontiguousarray(Ut, dtype=np.float64).reshape(self.B, 3)
        z = np.ascontiguousarray(matchLocation, dtype=np.float64).reshape(self.B, self.L, 2)
        m = np.ascontiguousarray(isMatching, dtype=np.uint8).reshape(self.B, self.L)
        capi.check(self._lib.srukf_step(self._h, capi.ptr(u), capi.ptr(z), capi.ptr(m)))

    def get_x_async(self, out_host_ptr: int):
        """Read-back of m_X_k that overlaps the next frame (srukf_get_x_async): `out_host_ptr` is the address of a
        pinned [B, n] float64 host buffer; complete after sync()."""
        capi.check(self._lib.srukf_get_x_async(self._h, int(out_host_ptr)))

    def GSLCholeskyUpdate(self, u: np.ndarray, flag4UpOrDown: int, flag4Order: int, n_new: int = 0):
        """CSLAM::GSLCholeskyUpdate (SLAM.h:347, SLAM.cpp:2106-2155) on every filter's m_S_k: u [B, n, k];
        flag values as in the reference (FLAG_4_UPDATING 0 / DOWNDATING 1, NEED_REORDER 0 / NEEDNOT_REORDER 1)."""
        u = np.ascontiguousarray(u, dtype=np.float64).reshape(self.B, self.n, -1)
        capi.check(self._lib.srukf_cholesky_update(self._h, capi.ptr(u), u.shape[2], int(flag4UpOrDown), int(flag4Order),
                                                   int(n_new)))

    def SLAM_dev(self, d_u: int, d_z: int, d_matched: int):
        """Same, inputs already in HBM (integer device addresses); asynchronous on the handle's own non-blocking
        stream: the buffers must be complete before the call (synchronise the stream that produced them)."""
        capi.check(self._lib.srukf_step_dev(self._h, d_u, d_z, d_matched))

    # ---- diagnostics --------------------------------------------------------------------------
    def flags(self) -> np.ndarray:
        f = np.empty(self.B, dtype=np.uint32)
        capi.check(self._lib.srukf_get_flags(self._h, capi.ptr(f)))
        return f

    def clear_flags(self):
        capi.check(self._lib.srukf_clear_flags(self._h))

    def stats(self, truth: np.ndarray) -> np.ndarray:
        truth = np.ascontiguousarray(truth, dtype=np.float64).reshape(self.B, 3)
        out = np.empty(8)
        capi.check(self._lib.srukf_stats(self._h, capi.ptr(truth), capi.ptr(out)))
        return out

    def sync(self):
        capi.check(self._lib.srukf_sync(self._h))

    def stream(self) -> int:
        s = C.c_uint64()
        capi.check(self._lib.srukf_stream(self._h, C.byref(s)))
        return s.value

    def launch_count(self) -> int:
        c = C.c_uint64()
        capi.check(self._lib.srukf_launch_count(self._h, C.byref(c)))
        return c.value

    def set_profiling(self, on: bool):
        capi.check(self._lib.srukf_set_profiling(self._h, 1 if on else 0))

    def kernel_times(self):
        """(ms[3], launches[3]) of k_predict / k_gain / k_downdate since profiling was switched on."""
        ms = np.zeros(3)
        cnt = np.zeros(3, dtype=np.uint64)
        capi.check(self._lib.srukf_get_kernel_times(self._h, capi.ptr(ms), capi.ptr(cnt)))
        return ms, cnt

    def get_x(self, out: np.ndarray | None = None) -> np.ndarray:
        """m_X_k only (the per-frame result a caller reads back)."""
        x = out if out is not None else np.empty((self.B, self.n))
        capi.check(self._lib.srukf_get_state(self._h, capi.ptr(x), None))
        return x

    def set_state_dev(self, b0: int, nb: int, d_x: int | None, d_S_packed: int | None):
        """Device-to-device load of filters [b0, b0+nb) (integer device addresses)."""
        capi.check(self._lib.srukf_set_state_dev(self._h, b0, nb, d_x, d_S_packed))


# ---- CSLAM helper methods that do not need a filter handle (SLAM.h:322,341,348,355) -------------------------
FLAG_4_UPDATING, FLAG_4_DOWNDATING = 0, 1            # SLAM.cpp:31-32
FLAG_4_NEED_REORDER, FLAG_4_NEEDNOT_REORDER = 0, 1   # SLAM.cpp:36-37


def modifiedCholeskyDecomposition(Cov: np.ndarray, epsilon: float = 1e-13, device: int = 0, return_flags: bool = False):
    """CSLAM::modifiedCholeskyDecomposition (SLAM.cpp:2197-2327) on the GPU: Cov [nb, n, n] or [n, n] -> sr (upper)."""
    G = np.ascontiguousarray(Cov, dtype=np.float64)
    single = G.ndim == 2
    G = G.reshape(-1, G.shape[-1], G.shape[-1])
    S = np.empty_like(G)
    fl = np.zeros(G.shape[0], dtype=np.uint32)
    capi.check(capi.load_library().srukf_mchol(device, G.shape[0], G.shape[-1], float(epsilon), capi.ptr(G), capi.ptr(S),
                                               capi.ptr(fl)))
    S = S[0] if single else S
    return (S, fl) if return_flags else S


def GSLQrDecomposition(A: np.ndarray, device: int = 0) -> np.ndarray:
    """CSLAM::GSLQrDecomposition (SLAM.cpp:2330-2353) on the GPU: A [nb, m, n] or [m, n] (m >= n) -> triu(R)."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    single = A.ndim == 2
    A = A.reshape(-1, A.shape[-2], A.shape[-1])
    R = np.empty((A.shape[0], A.shape[2], A.shape[2]))
    capi.check(capi.load_library().srukf_qr_R(device, A.shape[0], A.shape[1], A.shape[2], capi.ptr(A), capi.ptr(R)))
    return R[0] if single else R


def generateSigmaPoints(mu: np.ndarray, sr: np.ndarray, gamma: float, device: int = 0) -> np.ndarray:
    """CSLAM::generateSigmaPoints (SLAM.cpp:1148-1162) on the GPU: mu [nb, Na] or [Na], sr [nb, Na, Na] -> sigma
    [nb, Na, 2Na+1]."""
    mu = np.ascontiguousarray(mu, dtype=np.float64)
    single = mu.ndim == 1
    mu = mu.reshape(-1, mu.shape[-1])
    Na = mu.shape[1]
    sr = np.ascontiguousarray(sr, dtype=np.float64).reshape(mu.shape[0], Na, Na)
    out = np.empty((mu.shape[0], Na, 2 * Na + 1))
    capi.check(capi.load_library().srukf_generate_sigma_points(device, mu.shape[0], Na, float(gamma), capi.ptr(mu),
                                                               capi.ptr(sr), capi.ptr(out)))
    return out[0] if single else out


def fp64_peak_tflops(device: int = 0) -> float:
    v = C.c_double()
    capi.check(capi.load_library().srukf_fp64_peak(device, C.byref(v)))
    return v.value
