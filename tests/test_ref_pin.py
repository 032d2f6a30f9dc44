"""Pins oracle/srukf_oracle.c to the REFERENCE'S OWN TEXT.

oracle/_ref/libsrukf_ref.so holds the bodies of the CSLAM member functions on the hot path, copied verbatim from
/root/reference/MonoSLAM/SLAM.cpp at build time (oracle/ref_shim/extract_ref.py; nothing of the reference is committed)
and compiled against oracle/ref_shim/ref_shim.h (cv::Mat subset, MFC stand-ins; the GSL QR boundary is the restated
oracle_qr_decomp, cross-checked against LAPACK in test_oracle.py).  Every test here runs the reference's code and the
oracle on the same inputs.  Integer / flag outputs must be equal; floating-point outputs are required to be BIT-IDENTICAL
where the oracle follows the reference's operation order (it does, everywhere below), which is stronger than the
1e-13 the plan asked for.
"""
import glob
import os

import numpy as np
import pytest

import synth
from conftest import relmax

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# SURVEY.md 8(a) cites these ranges; the extractor finds the functions by name and must land on the same lines
CITED = {
    "calculateSampleParameter": (1050, 1103), "expandMatrix": (1123, 1135), "generateSigmaPoints": (1148, 1162),
    "passSigmaThroughMotionFunction": (1476, 1532), "QrAndCholeskyForMotion": (1539, 1595),
    "predictMeasurement": (1604, 1608), "passSigmaThroughMesaurementFunction": (1615, 1691),
    "QrAndCholeskyForMeasurement": (1700, 1748), "calculateOneFeatureCrossCovariance": (2020, 2038),
    "KalmanUpdate": (2048, 2104), "GSLCholeskyUpdate": (2106, 2155), "modifiedCholeskyDecomposition": (2197, 2327),
    "GSLQrDecomposition": (2330, 2353), "deleteOneFeature": (2637, 2706), "distortOnePointRW": (3177, 3213),
    "coordinatesState2World": (3250, 3276), "coordinatesCamera2Image": (3324, 3347), "predictMotion": (1343, 1466),
}


def run_frame(r, f, u_seed, z, matched, n_new=0):
    """one predictMotion / predictMeasurement / KalmanUpdate frame on both; the oracle gets the control the reference
    derived from the odometry poses (SLAM.cpp:1446-1450)"""
    import ref as R
    u, mt = r.predict_motion_odometry(*R.control_to_odometry(u_seed))
    assert np.abs(u - u_seed).max() < 1e-15
    f.predict_motion(u)
    r.predict_measurement()
    f.predict_measurement()
    if n_new:
        f.set_new_features(n_new)
    r.kalman_update(z, matched, n_new)
    f.kalman_update(z, matched)
    return u


def assert_same_state(r, f, exact=True):
    xr, Sr = r.get_state()
    xo, So = f.get_state()
    if exact:
        assert np.array_equal(xo, xr) and np.array_equal(So, Sr)
    else:
        assert relmax(xo, xr) < 1e-13 and relmax(So.T @ So, Sr.T @ Sr) < 1e-13


def test_extracted_text_is_the_cited_text(reference):
    got = {name: (a, b) for name, a, b, _ in reference.manifest()}
    for name, rng in CITED.items():
        assert got[name] == rng, (name, got[name], rng)


def test_reference_defaults_equal_oracle_defaults(reference, oracle):
    rp = reference.Slam().params()
    op = oracle.default_params()
    for k in ("cam_dx", "cam_dy", "cam_cx", "cam_cy", "cam_k1", "cam_k2", "cam_f", "a1", "a2", "a3", "a4",
              "sigma_measure", "epsilon", "alpha", "beta"):
        assert rp[k] == getattr(op, k), k
    assert (rp["image_width"], rp["image_height"], rp["weight_type"]) == (op.image_width, op.image_height, op.weight_type)
    n = synth.Noise()
    assert (rp["rho"], rp["sigma_rho"]) == (n.rho0, n.sigma_rho)
    assert [rp["sigma_x"], rp["sigma_y"], rp["sigma_z"], rp["sigma_theta"]] == list(n.S4)


@pytest.mark.parametrize("wt", [0, 1, 2])
def test_sample_parameters(reference, oracle, wt):
    r = reference.Slam(wt)
    for Na in (9, 69, 129, 309, 405, 1209):
        assert r.sample_parameters(Na) == oracle.sample_parameters(Na, oracle.default_params(weight_type=wt))


def test_modified_cholesky(reference, oracle):
    rng = np.random.default_rng(7)
    r = reference.Slam()
    for n in (1, 2, 7, 33, 64):
        A = rng.standard_normal((n, n))
        cases = [A @ A.T + n * np.eye(n), A + A.T, A[:, :max(1, n // 2)] @ A[:, :max(1, n // 2)].T,
                 -(A @ A.T), np.zeros((n, n)), np.diag(rng.standard_normal(n))]
        for G in cases:
            S, E, nmod = oracle.mchol(G)
            assert np.array_equal(S, r.mchol(G))


def test_camera_chain(reference, oracle):
    rng = np.random.default_rng(11)
    p = oracle.default_params()
    r = reference.Slam()
    for _ in range(200):
        ux, uy = rng.uniform(-100, 800), rng.uniform(-100, 600)
        assert r.distort(ux, uy) == oracle.distort(p, ux, uy)
        assert r.undistort(ux, uy) == oracle.undistort(p, ux, uy)
        feat = np.array([rng.normal(0, 1), rng.normal(0, 1), 0.0, rng.uniform(-.3, .3), rng.uniform(-.3, .3), rng.uniform(.1, .6)])
        pos = np.array([rng.normal(0, .5), rng.normal(0, .5), 0.0])
        th, err = rng.uniform(-np.pi, np.pi), rng.normal(0, 3, 2)
        assert r.project(feat, pos, th, err) == oracle.project(p, feat, pos, th, err)
    # second distortion coefficient (cam_k2 != 0 is not the default but the code handles it, SLAM.cpp:3190-3205)
    r.set_camera(1e-4, 3e-9)
    p2 = oracle.default_params(cam_k2=3e-9)
    for _ in range(50):
        ux, uy = rng.uniform(0, 640), rng.uniform(0, 480)
        assert r.distort(ux, uy) == oracle.distort(p2, ux, uy)


@pytest.mark.parametrize("L,steps,match_prob", [(1, 5, 1.0), (3, 6, 1.0), (8, 5, 0.6), (20, 3, 1.0)])
def test_whole_frames_are_bit_identical(reference, oracle, L, steps, match_prob):
    sc = synth.make_scenario(L, 2, steps, match_prob=match_prob)
    for b in range(2):
        r = reference.Slam()
        r.set_state(sc.x0[b], sc.S0[b])
        f = oracle.Filter(L, oracle.default_params(downdate_mode=0))
        f.set_state(sc.x0[b], sc.S0[b])
        for s in range(steps):
            u, _ = r.predict_motion_odometry(*reference.control_to_odometry(sc.u[s, b]))
            f.predict_motion(u)
            assert_same_state(r, f)                       # a2-a5: motion + QR
            r.predict_measurement()
            f.predict_measurement()
            ho, sio, vo = f.prediction()
            hr, sir, vr = r.prediction()
            assert np.array_equal(vo, vr) and np.array_equal(ho, hr) and np.array_equal(sio, sir)   # a6, a7
            r.kalman_update(sc.z[s, b], sc.matched[s, b])
            f.kalman_update(sc.z[s, b], sc.matched[s, b])
            assert_same_state(r, f)                       # a8-a11


def test_headline_size_one_frame(reference, oracle):
    """L = 50 (n = 304, 619 sigma points): 100 dense S^T S + modified-Cholesky passes of the reference's own code"""
    sc = synth.make_scenario(50, 1, 1)
    r = reference.Slam()
    r.set_state(sc.x0[0], sc.S0[0])
    f = oracle.Filter(50, oracle.default_params(downdate_mode=0))
    f.set_state(sc.x0[0], sc.S0[0])
    run_frame(r, f, sc.u[0, 0], sc.z[0, 0], sc.matched[0, 0])
    assert_same_state(r, f)


@pytest.mark.parametrize("wt", [1, 2])
def test_other_weight_types(reference, oracle, wt):
    sc = synth.make_scenario(4, 1, 3)
    r = reference.Slam(wt)
    r.set_state(sc.x0[0], sc.S0[0])
    f = oracle.Filter(4, oracle.default_params(downdate_mode=0, weight_type=wt))
    f.set_state(sc.x0[0], sc.S0[0])
    for s in range(3):
        run_frame(r, f, sc.u[s, 0], sc.z[s, 0], sc.matched[s, 0])
        assert_same_state(r, f)


def test_no_matches_returns_early(reference, oracle):
    sc = synth.make_scenario(3, 1, 1)
    r = reference.Slam()
    r.set_state(sc.x0[0], sc.S0[0])
    f = oracle.Filter(3, oracle.default_params(downdate_mode=0))
    f.set_state(sc.x0[0], sc.S0[0])
    run_frame(r, f, sc.u[0, 0], sc.z[0, 0], np.zeros(3, dtype=np.uint8))
    assert_same_state(r, f)


def test_oracle_modes_1_and_2_against_the_reference(reference, oracle):
    """mode 2 (dense product) is the same bits; mode 1 (carry-P) is the same mathematics: <= 1e-12 on P"""
    L, steps = 6, 4
    sc = synth.make_scenario(L, 1, steps)
    r = reference.Slam()
    r.set_state(sc.x0[0], sc.S0[0])
    fs = {m: oracle.Filter(L, oracle.default_params(downdate_mode=m)) for m in (1, 2)}
    for f in fs.values():
        f.set_state(sc.x0[0], sc.S0[0])
    for s in range(steps):
        u, _ = r.predict_motion_odometry(*reference.control_to_odometry(sc.u[s, 0]))
        r.predict_measurement()
        r.kalman_update(sc.z[s, 0], sc.matched[s, 0])
        for f in fs.values():
            f.step(u, sc.z[s, 0], sc.matched[s, 0])
    xr, Sr = r.get_state()
    x2, S2 = fs[2].get_state()
    assert np.array_equal(x2, xr) and np.array_equal(S2, Sr)
    x1, S1 = fs[1].get_state()
    assert relmax(x1, xr) < 1e-12 and relmax(S1.T @ S1, Sr.T @ Sr) < 1e-12


def _keypoints(rng, M):
    cam = synth.Camera()
    rad, ang = rng.uniform(30, 150, M), rng.uniform(0, 2 * np.pi, M)
    kp = np.stack([cam.cx + rad * np.cos(ang), cam.cy + rad * np.sin(ang)], -1)
    return kp.astype(np.float32).astype(np.float64)     # KeyPoint::pt is Point2f in the reference


@pytest.mark.parametrize("M", [1, 4, 9])
def test_map_life_cycle(reference, oracle, M):
    """f1 / f2: initialisation at frame 1, a NEED_REORDER update on the frame that added features, augmentation of a
    non-empty map, deletion of every position"""
    rng = np.random.default_rng(100 + M)
    p = oracle.default_params()
    nz = synth.Noise()
    x4, S4 = np.array([0.05, -0.1, 0.0, 0.6]), np.diag(nz.S4)
    kp = _keypoints(rng, M)
    r = reference.Slam()
    r.set_state(x4, S4)
    r.init_features(kp, nz.rho0, nz.sigma_rho)
    x, S = oracle.init_features(p, x4, S4, kp, nz.rho0, nz.sigma_rho)
    xr, Sr = r.get_state()
    assert np.array_equal(x, xr) and np.array_equal(S, Sr)
    # the frame that follows an addition updates through the pivoted branch (m_nAddings != 0, SLAM.cpp:2082-2085)
    cam = synth.Camera()
    f = oracle.Filter(M, oracle.default_params(downdate_mode=0))
    f.set_state(x, S)
    u = np.array([0.025, 0.005, 0.025])
    uu, _ = r.predict_motion_odometry(*reference.control_to_odometry(u))
    f.predict_motion(uu)
    r.predict_measurement()
    f.predict_measurement()
    hbar, _, vis = r.prediction()
    z = hbar + rng.normal(0, 1, hbar.shape)
    f.set_new_features(M)
    r.kalman_update(z, vis, n_new=M)
    f.kalman_update(z, vis)
    assert_same_state(r, f)
    x, S = f.get_state()
    # augmentation of the non-empty map
    kp2 = _keypoints(rng, 2)
    r.init_features(kp2, nz.rho0, nz.sigma_rho)
    x2, S2 = oracle.add_features(p, x, S, kp2, nz.rho0, nz.sigma_rho)
    xr, Sr = r.get_state()
    assert np.array_equal(x2, xr) and np.array_equal(S2, Sr)
    # deletion at every position (first, middle, last take different branches, SLAM.cpp:2643-2662)
    for id_ in range(M + 2):
        rr = reference.Slam()
        rr.set_state(x2, S2)
        rr.delete_feature(id_)
        xd, Sd = oracle.delete_feature(p, x2, S2, id_)
        xr, Sr = rr.get_state()
        assert np.array_equal(xd, xr) and np.array_equal(Sd, Sr)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "*.npz"))))
def test_reference_reproduces_the_committed_fixtures(reference, path):
    """tests/golden/*.npz were written by tests/golden/make_golden.py from THIS library (the reference's own code);
    re-running it must give the committed numbers to the bit"""
    g = np.load(path)
    if "source" not in g.files:
        pytest.skip("fixture predates the reference-derived generator")
    L, B, steps = int(g["L"]), int(g["B"]), int(g["steps"])
    wt = int(g["weight_type"]) if "weight_type" in g.files else 0
    for b in range(B):
        r = reference.Slam(wt)
        r.set_state(g["x0"][b], g["S0"][b])
        for s in range(steps):
            u, _ = r.predict_motion_odometry(g["odo"][s, b, 0], g["odo"][s, b, 1])
            assert np.array_equal(u, g["u"][s, b])
            r.predict_measurement()
            r.kalman_update(g["z"][s, b], g["matched"][s, b])
            x, S = r.get_state()
            assert np.array_equal(x, g["x"][s, b]) and np.array_equal(S.T @ S, g["P"][s, b])
