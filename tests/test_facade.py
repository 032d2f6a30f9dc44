"""The header-only C++ facade (include/SLAM.h, the reference's CSLAM method names) builds against the C ABI;
on a GPU box it must reproduce the CPU oracle (frames and helper methods), without a GPU it must fail loudly."""
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


def helper_inputs(n, seed=5):
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((n, n))
    G = M @ M.T + 0.5 * np.eye(n)
    A = rng.standard_normal((2 * n, n))
    return G, A


def write_scenario(path, sc):
    n = 6 * sc.L + 4
    G, A = helper_inputs(n)
    with open(path, "wb") as f:
        f.write(struct.pack("ii", sc.L, sc.steps))
        f.write(np.ascontiguousarray(sc.x0[0]).tobytes())
        f.write(np.ascontiguousarray(sc.S0[0]).tobytes())
        f.write(np.ascontiguousarray(sc.u[:, 0]).tobytes())
        f.write(np.ascontiguousarray(sc.z[:, 0]).tobytes())
        f.write(np.ascontiguousarray(G).tobytes())
        f.write(np.ascontiguousarray(A).tobytes())


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
def test_facade_matches_the_oracle(tmp_path, built_lib, oracle):
    """The C++ facade (reference method names, data through members) against the CPU oracle: whole frames through the
    no-argument SLAM() (chi-square gate of dataAssociation decides the matches) and through a caller-supplied
    association, then the helper methods modifiedCholeskyDecomposition / GSLQrDecomposition / generateSigmaPoints /
    GSLCholeskyUpdate."""
    from conftest import relmax
    exe = build_facade(tmp_path, built_lib)
    L, steps = 6, 4
    n = 6 * L + 4
    sc = synth.make_scenario(L, 1, steps)
    write_scenario(tmp_path / "sc.bin", sc)
    out = subprocess.run([exe, str(tmp_path / "sc.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    vals = [float(v) for v in out.stdout.split()]
    raw = np.fromfile(tmp_path / "out.bin")
    sizes = [n, n * n, 1, n * n, n * n, n * (2 * n + 1), n * n]
    assert raw.size == sum(sizes)
    xf, Sf, tr, srf, Rf, sigf, S2f = np.split(raw, np.cumsum(sizes)[:-1])
    Sf, srf, Rf, S2f = (a.reshape(n, n) for a in (Sf, srf, Rf, S2f))
    # ---- frames: the oracle with the same association rule (gate on even frames, "all visible" on odd ones)
    f = oracle.Filter(L, oracle.default_params(downdate_mode=0))
    f.set_state(sc.x0[0], sc.S0[0])
    nm = 0
    for s in range(steps):
        f.predict_motion(sc.u[s, 0])
        f.predict_measurement()
        _, _, vis = f.prediction()
        m = f.chi2_gate(sc.z[s, 0])[0] if s % 2 == 0 else vis
        nm = int(m.sum())
        f.kalman_update(sc.z[s, 0], m)
    xo, So = f.get_state()
    assert relmax(xf, xo) < 1e-9 and relmax(Sf.T @ Sf, So.T @ So) < 1e-9
    assert np.array_equal(np.array(vals[:4]), xf[n - 4:])
    assert vals[4] == pytest.approx(np.trace(So.T @ So), rel=1e-9)
    assert int(vals[5]) == nm and int(vals[6]) == L
    # ---- helpers
    G, A = helper_inputs(n)
    So_m, _, _ = oracle.mchol(G)
    assert relmax(srf, So_m) < 1e-11
    assert relmax(Rf, oracle.qr_R(A)) < 1e-11
    gam = oracle.sample_parameters(n)["gamma"]
    sig = np.empty((n, 2 * n + 1))
    sig[:, 0] = xf
    for i in range(n):
        sig[:, 1 + i] = xf * 1 + Sf[i] * gam + 0
        sig[:, 1 + n + i] = xf * 1 + Sf[i] * ((-1) * gam) + 0
    assert np.array_equal(sigf.reshape(n, 2 * n + 1), sig)
    S2 = Sf.copy()
    for u in (0.25 * Sf[0], 0.25 * Sf[n - 4]):       # SLAM.cpp:2116-2153, NEEDNOT_REORDER / DOWNDATING
        S2, _, _ = oracle.mchol(S2.T @ S2 - np.outer(u, u))
    assert relmax(S2f.T @ S2f, S2.T @ S2) < 1e-10
