"""oracle/ref_shim/ref_shim.h restates the OpenCV 2.4 primitives the reference's SRUKF functions call.  These tests check
that restatement against the cv2 (4.x) Python bindings, whose CV_64F arithmetic for these functions is the same:
closed-form inverses, addWeighted's a*alpha + b*beta + gamma, divide's 0 on a zero divisor, minMaxLoc's first
occurrence in row-major order, and the header / ROI write-through semantics."""
import ctypes as C

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def shim(reference):
    L = reference.lib()
    L.shim_gemm.argtypes = [_dp, C.c_int, C.c_int, _dp, C.c_int, _dp]
    L.shim_inv.argtypes = [_dp, C.c_int, _dp]
    L.shim_add_weighted.argtypes = [_dp, C.c_double, _dp, C.c_double, C.c_double, C.c_int, C.c_int, _dp]
    L.shim_divide.argtypes = [_dp, _dp, C.c_int, C.c_int, _dp]
    L.shim_min_max_loc.argtypes = [_dp, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                   np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")]
    L.shim_roi_assign.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]
    L.shim_roi_shift_left.argtypes = [_dp, C.c_int, C.c_int, C.c_int]
    return L


def test_gemm(shim):
    rng = np.random.default_rng(0)
    for m, k, n in ((1, 1, 1), (3, 1, 2), (7, 5, 3), (40, 33, 17), (124, 124, 124)):
        A, B = rng.standard_normal((m, k)), rng.standard_normal((k, n))
        out = np.zeros((m, n))
        shim.shim_gemm(A, m, k, B, n, out)
        # one accumulator per element, k ascending: this is numpy's float64 loop below, bit for bit
        ref = np.zeros((m, n))
        for kk in range(k):
            ref += np.outer(A[:, kk], B[kk])
        assert np.array_equal(out, ref)
        got = cv2.gemm(A, B, 1.0, None, 0.0)
        assert np.abs(out - got).max() <= 1e-14 * max(1.0, np.abs(got).max()) * k


def test_small_inverses_are_the_closed_forms(shim):
    rng = np.random.default_rng(1)
    for n in (1, 2, 3):
        for _ in range(50):
            A = rng.standard_normal((n, n))
            out = np.zeros((n, n))
            shim.shim_inv(A, n, out)
            ok, got = cv2.invert(A, flags=cv2.DECOMP_LU)
            assert np.array_equal(out, got), (n, out, got)
    for n in (2, 3):                                   # singular -> zeros (cv::invert returns 0 and clears dst)
        A = np.ones((n, n))
        out = np.full((n, n), 7.0)
        shim.shim_inv(A, n, out)
        assert not out.any()
    for n in (4, 6, 12):                               # LU with partial pivoting
        A = rng.standard_normal((n, n)) + n * np.eye(n)
        out = np.zeros((n, n))
        shim.shim_inv(A, n, out)
        ok, got = cv2.invert(A, flags=cv2.DECOMP_LU)
        assert np.abs(out - got).max() <= 1e-13 * np.abs(got).max()


def test_add_weighted_divide_minmaxloc(shim):
    rng = np.random.default_rng(2)
    A, B = rng.standard_normal((9, 5)), rng.standard_normal((9, 5))
    out = np.zeros_like(A)
    shim.shim_add_weighted(A, 0.7, B, -1.3, 0.25, 9, 5, out)
    assert np.array_equal(out, A * 0.7 + B * -1.3 + 0.25)      # 2.4's addWeighted64f: src1*alpha + src2*beta + gamma
    # cv2 4.x vectorises this with fused multiply-adds, so it agrees to an ulp, not to the bit; the beta = 0 form the
    # reference uses for the central sigma point is bit-identical
    assert np.abs(out - cv2.addWeighted(A, 0.7, B, -1.3, 0.25)).max() <= 4.5e-16 * np.abs(out).max()
    shim.shim_add_weighted(A, 0.7, B, 0.0, 0.0, 9, 5, out)
    assert np.array_equal(out, cv2.addWeighted(A, 0.7, B, 0.0, 0.0))
    B[2, 3] = 0.0
    B[0, 0] = 0.0
    shim.shim_divide(A, B, 9, 5, out)
    nz = B != 0
    assert np.array_equal(out[nz], cv2.divide(A, B)[nz])
    # OpenCV 2.4's div_ returns 0 where the divisor is 0 (4.x switched floating-point division to IEEE inf / nan);
    # the reference was built against 2.4.3, and modifiedCholeskyDecomposition relies on the 0 (SLAM.cpp:2213-2216)
    assert out[2, 3] == 0.0 and out[0, 0] == 0.0
    A[4, 1] = A[7, 2] = 9.0                            # a tie: the first occurrence in row-major order wins
    A[1, 4] = A[3, 0] = -9.0
    mn, mx = C.c_double(), C.c_double()
    loc = np.zeros(4, dtype=np.int32)
    shim.shim_min_max_loc(A, 9, 5, C.byref(mn), C.byref(mx), loc)
    cmn, cmx, cmnl, cmxl = cv2.minMaxLoc(A)
    assert (mn.value, mx.value) == (cmn, cmx)
    assert tuple(loc[:2]) == cmnl and tuple(loc[2:]) == cmxl


def test_roi_headers_write_through(shim):
    A = np.arange(30, dtype=np.float64).reshape(5, 6)
    B = A.copy()
    shim.shim_roi_assign(B, 5, 6, 1, 3, 2, 5, 2.0)
    E = A.copy()
    E[1:3, 2:5] *= 2.0
    assert np.array_equal(B, E)
    B = A.copy()
    shim.shim_roi_shift_left(B, 5, 6, 2)               # deleteOneFeature's overlapping block moves, SLAM.cpp:2650-2658
    E = A.copy()
    E[:, 0:4] = A[:, 2:6]
    assert np.array_equal(B, E)
