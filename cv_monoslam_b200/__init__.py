"""cv_monoslam_b200 -- B200-native (sm_100a, FP64) batched SRUKF predict/update.

A from-scratch implementation of the one data-parallel hot path of junliu111/CV-MonoSLAM
(MonoSLAM/SLAM.cpp: predictMotion / predictMeasurement / KalmanUpdate), batched over independent
filters.  The compute path is hand-written CUDA behind a C ABI (include/srukf.h, libsrukf_b200.so);
this package is the thin host-side mirror of the reference's CSLAM interface used by tests and
bench.py.  There is no CPU fallback: importing works anywhere, but creating a filter batch without
the built library or without a CUDA device raises.
"""
from .capi import SrukfError, SrukfParams, default_params, lib_path, load_library  # noqa: F401
from .slam import (CSLAMBatch, GSLQrDecomposition, generateSigmaPoints,  # noqa: F401
                   modifiedCholeskyDecomposition)

__all__ = ["CSLAMBatch", "modifiedCholeskyDecomposition", "GSLQrDecomposition", "generateSigmaPoints", "SrukfParams", "SrukfError", "default_params", "load_library", "lib_path"]
