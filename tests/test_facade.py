"""The header-only C++ facade (include/SLAM.h, the reference's CSLAM method names) builds against the C ABI;
on a GPU box it must reproduce the Python/ctypes path bit for bit, without a GPU it must fail loudly."""
import os
import struct
import subprocess

import numpy as np
import pytest

import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_facade(tmp_path, built_lib):
    from cv_monoslam_b200 import capi
    exe = str(tmp_path / "test_facade")
    libdir = os.path.dirname(capi.lib_path())
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_facade.cpp"),
           "-o", exe, "-L", libdir, "-l:libsrukf_b200.so", "-Wl,-rpath," + libdir]
    subprocess.check_call(cmd)
    return exe


def write_scenario(path, sc):
    with open(path, "wb") as f:
        f.write(struct.pack("ii", sc.L, sc.steps))
        f.write(np.ascontiguousarray(sc.x0[0]).tobytes())
        f.write(np.ascontiguousarray(sc.S0[0]).tobytes())
        f.write(np.ascontiguousarray(sc.u[:, 0]).tobytes())
        f.write(np.ascontiguousarray(sc.z[:, 0]).tobytes())


def test_facade_compiles_and_refuses_to_run_without_gpu(tmp_path, built_lib):
    import torch
    exe = build_facade(tmp_path, built_lib)
    sc = synth.make_scenario(3, 1, 2)
    write_scenario(tmp_path / "sc.bin", sc)
    out = subprocess.run([exe, str(tmp_path / "sc.bin")], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert out.returncode == 0, out.stderr
    else:
        assert out.returncode == 3 and "no CUDA device" in out.stderr


@pytest.mark.gpu
def test_facade_matches_ctypes_path(tmp_path, built_lib):
    from cv_monoslam_b200 import CSLAMBatch
    exe = build_facade(tmp_path, built_lib)
    L, steps = 6, 4
    sc = synth.make_scenario(L, 1, steps)
    write_scenario(tmp_path / "sc.bin", sc)
    out = subprocess.run([exe, str(tmp_path / "sc.bin")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    vals = [float(v) for v in out.stdout.split()]
    g = CSLAMBatch(1, L)
    g.set_state(sc.x0, sc.S0)
    for s in range(steps):
        g.predictMotion(sc.u[s])
        g.predictMeasurement()
        g.KalmanUpdate(sc.z[s], sc.matched[s])
    x, S = g.get_state()
    n = 6 * L + 4
    assert np.array_equal(np.array(vals[:4]), x[0, n - 4:])
    assert vals[4] == pytest.approx(np.trace(S[0].T @ S[0]), rel=1e-13)
    assert int(vals[5]) == L and int(vals[6]) == L
