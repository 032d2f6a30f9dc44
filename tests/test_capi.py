"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/srukf.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "srukf.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(srukf_[a-z_]+)\s*\(", txt)))


def test_header_declares_the_path_entry_points():
    syms = header_symbols()
    for s in ("srukf_create", "srukf_destroy", "srukf_set_state", "srukf_get_state", "srukf_predict_motion",
              "srukf_predict_measurement", "srukf_kalman_update", "srukf_step", "srukf_get_cov_block",
              "srukf_get_flags", "srukf_stats", "srukf_sync"):
        assert s in syms


def test_library_exports_every_header_symbol(built_lib):
    from cv_monoslam_b200 import capi
    for s in header_symbols():
        assert hasattr(built_lib, s), f"{s} declared in include/srukf.h but not exported"
        assert s in capi.SYMBOLS, f"{s} has no ctypes prototype in capi.py"
    assert b"sm_100a" in built_lib.srukf_version()


def test_library_is_built_for_sm_100a_only(built_lib):
    import shutil
    import subprocess
    from cv_monoslam_b200 import capi
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", capi.lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_params_struct_matches_defaults(built_lib):
    from cv_monoslam_b200 import capi
    p = capi.default_params()
    assert (p.cam_cx, p.cam_cy, p.cam_f) == (310.1129, 236.7526, 2.1735)   # SLAM.cpp:331-335
    assert (p.a1, p.a2, p.a3, p.a4) == (8, 8, 8, 8)                         # :195-198
    assert p.sigma_measure == 3.0 and p.epsilon == 1e-13 and p.newton_iters == 100
    assert p.weight_type == 0 and p.downdate_mode == 0
    assert C.sizeof(capi.SrukfParams) == 144


def test_bad_arguments_are_rejected_without_touching_a_device(built_lib):
    from cv_monoslam_b200 import capi
    h = C.c_void_p()
    p = capi.default_params()
    assert built_lib.srukf_create(0, 0, 5, C.byref(p), C.byref(h)) == capi.SRUKF_EINVAL
    assert built_lib.srukf_create(0, 4, 0, C.byref(p), C.byref(h)) == capi.SRUKF_EINVAL
    assert built_lib.srukf_set_state(None, None, None) == capi.SRUKF_EINVAL
    assert built_lib.srukf_step(None, None, None, None) == capi.SRUKF_EINVAL
    assert built_lib.srukf_destroy(None) == capi.SRUKF_OK


def test_no_cpu_fallback(built_lib):
    """Without a CUDA device the product path must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from cv_monoslam_b200 import CSLAMBatch, SrukfError, capi
    with pytest.raises(SrukfError) as ei:
        CSLAMBatch(4, 3)
    assert ei.value.code == capi.SRUKF_ENODEV
    assert b"no CPU fallback" in built_lib.srukf_last_error()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cv_monoslam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "srukf_oracle" not in txt and "import oracle" not in txt, f
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def test_tri_pack_roundtrip():
    from cv_monoslam_b200.slam import tri_pack, tri_unpack
    rng = np.random.default_rng(0)
    S = np.triu(rng.standard_normal((3, 10, 10)))
    Sp = tri_pack(S)
    assert Sp.shape == (3, 55)
    assert np.array_equal(tri_unpack(Sp, 10), S)
    # layout contract of include/srukf.h: row i at offset i*n - i(i-1)/2
    n = 10
    for i in (0, 3, 9):
        off = i * n - i * (i - 1) // 2
        assert np.array_equal(Sp[0, off:off + n - i], S[0, i, i:])
