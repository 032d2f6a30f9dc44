"""GPU tests of the stand-alone helper entry points (CSLAM helper methods, SLAM.h:341,347-348,355) through the C ABI,
against the CPU oracle and -- where the prebuilt oracle/_ref library travelled to the box -- the reference's own code."""
import numpy as np
import pytest

from conftest import relmax
import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    from cv_monoslam_b200 import build, capi
    build.build()
    capi.load_library()
    import cv_monoslam_b200 as pkg
    return pkg


def test_modified_cholesky_matches_oracle(gpu, oracle):
    rng = np.random.default_rng(3)
    for n in (1, 5, 33, 124, 304):
        A = rng.standard_normal((n, n))
        pd = A @ A.T + n * np.eye(n)
        S = gpu.modifiedCholeskyDecomposition(pd)
        So, _, _ = oracle.mchol(pd)
        assert np.allclose(np.tril(S, -1), 0)
        assert relmax(S, So) < 1e-11, n
    # a batch with semi-definite and indefinite members: the pivots are modified (flags say so); P is compared
    n = 40
    A = rng.standard_normal((n, n))
    batch = np.stack([A @ A.T + n * np.eye(n), A[:, :20] @ A[:, :20].T, A + A.T, np.diag(rng.standard_normal(n))])
    S, fl = gpu.modifiedCholeskyDecomposition(batch, return_flags=True)
    # the semi-definite member is ill-posed for ANY implementation: once the rank is exhausted the pivots are rounding
    # noise, GMW's theta^2/beta^2 term amplifies it (the oracle itself moves by 1e-9 under a 1e-16 perturbation of G)
    tol = [1e-11, 1e-6, 1e-9, 1e-9]
    for i in range(4):
        So, E, nmod = oracle.mchol(batch[i])
        assert relmax(S[i].T @ S[i], So.T @ So) < tol[i], i
        assert bool(fl[i] & 6) == (nmod > 0)


def test_qr_matches_oracle(gpu, oracle):
    rng = np.random.default_rng(4)
    for m, n in ((2, 2), (7, 3), (40, 40), (618, 304), (258, 2)):
        A = rng.standard_normal((m, n))
        R = gpu.GSLQrDecomposition(A)
        Ro = oracle.qr_R(A)
        assert np.allclose(np.tril(R, -1), 0)
        assert relmax(R, Ro) < 1e-11, (m, n)
    A = rng.standard_normal((6, 4))
    A[1:, 1] = 0.0                      # a column whose sub-diagonal part is zero: tau = 0, column untouched
    assert relmax(gpu.GSLQrDecomposition(A), oracle.qr_R(A)) < 1e-12
    batch = rng.standard_normal((5, 30, 9))
    R = gpu.GSLQrDecomposition(batch)
    for i in range(5):
        assert relmax(R[i], oracle.qr_R(batch[i])) < 1e-11


def test_sigma_points_are_the_reference_formula(gpu):
    rng = np.random.default_rng(5)
    Na, gam = 29, 1.7320508075688772
    mu = rng.standard_normal((3, Na))
    sr = np.triu(rng.standard_normal((3, Na, Na)))
    sig = gpu.generateSigmaPoints(mu, sr, gam)
    for b in range(3):
        assert np.array_equal(sig[b, :, 0], mu[b])
        for i in range(Na):   # addWeighted(mu, 1, sr.row(i).t(), +-gamma, 0), SLAM.cpp:1159-1160
            assert np.array_equal(sig[b, :, 1 + i], mu[b] * 1 + sr[b, i] * gam + 0)
            assert np.array_equal(sig[b, :, 1 + Na + i], mu[b] * 1 + sr[b, i] * ((-1) * gam) + 0)


def _running_states(oracle, L, B, steps=2):
    sc = synth.make_scenario(L, B, steps)
    p = oracle.default_params(downdate_mode=0)
    x, S = sc.x0.copy(), sc.S0.copy()
    oracle.batch_step(p, x, S, sc.u, sc.z, sc.matched, 4)
    return x, S


@pytest.mark.parametrize("updown", [0, 1])
def test_cholesky_update_needs_no_reorder(gpu, oracle, updown):
    """GSLCholeskyUpdate(u, UPDATING / DOWNDATING, NEEDNOT_REORDER), SLAM.cpp:2139-2153, column by column"""
    L, B, k = 6, 3, 3
    x, S = _running_states(oracle, L, B)
    rng = np.random.default_rng(8)
    n = 6 * L + 4
    U = np.stack([(0.3 * S[b, rng.integers(0, n, k)]).T for b in range(B)])      # [B, n, k]: scaled rows of S keep P definite
    g = gpu.CSLAMBatch(B, L)
    g.set_state(x, S)
    g.GSLCholeskyUpdate(U, updown, 1)
    _, Sg = g.get_state()
    for b in range(B):
        So = S[b].copy()
        for c in range(k):
            u = U[b, :, c]
            So, _, _ = oracle.mchol(So.T @ So + (1.0 if updown == 0 else -1.0) * np.outer(u, u))
        assert relmax(Sg[b].T @ Sg[b], So.T @ So) < 1e-9
    g.close()


def test_cholesky_update_against_the_reference_code(gpu, oracle):
    """the same entry point against the reference's own GSLCholeskyUpdate (both orders), where oracle/_ref travelled"""
    import ref as R
    if not R.available():
        pytest.skip("oracle/_ref/libsrukf_ref.so is not on this box")
    L, B = 5, 2
    x, S = _running_states(oracle, L, B)
    n = 6 * L + 4
    U = np.stack([(0.3 * S[b, [1, n - 2]]).T for b in range(B)])
    for order, n_new in ((R.NEEDNOT_REORDER, 0), (R.NEED_REORDER, 2)):
        g = gpu.CSLAMBatch(B, L)
        g.set_state(x, S)
        g.GSLCholeskyUpdate(U, R.DOWNDATING, order, n_new)
        _, Sg = g.get_state()
        for b in range(B):
            Sr = R.Slam().cholesky_update(S[b], U[b], R.DOWNDATING, order, n_new=n_new, nmap=L)
            assert relmax(Sg[b].T @ Sg[b], Sr.T @ Sr) < 1e-9, (order, b)
        g.close()


def test_step_after_helper_update_uses_the_new_factor(gpu, oracle):
    """the carried covariance of the fused path is rebuilt by the helper: a following frame matches the oracle"""
    L, B = 4, 2
    sc = synth.make_scenario(L, B, 3)
    n = 6 * L + 4
    g = gpu.CSLAMBatch(B, L)
    g.set_state(sc.x0, sc.S0)
    U = np.stack([(0.2 * sc.S0[b, [0, n - 1]]).T for b in range(B)])
    g.GSLCholeskyUpdate(U, 1, 1)
    g.SLAM(sc.u[0], sc.z[0], sc.matched[0])
    xg, Sg = g.get_state()
    for b in range(B):
        So = sc.S0[b].copy()
        for c in range(2):
            So, _, _ = oracle.mchol(So.T @ So - np.outer(U[b, :, c], U[b, :, c]))
        f = oracle.Filter(L, oracle.default_params(downdate_mode=0))
        f.set_state(sc.x0[b], So)
        f.step(sc.u[0, b], sc.z[0, b], sc.matched[0, b])
        xo, S1 = f.get_state()
        assert relmax(xg[b], xo) < 1e-9 and relmax(Sg[b].T @ Sg[b], S1.T @ S1) < 1e-9
    g.close()


def test_async_readback_and_double_buffered_inputs(gpu):
    """srukf_step's alternating input buffers + srukf_get_x_async give the same bits as the synchronous path"""
    import torch
    L, B, steps = 5, 64, 6
    sc = synth.make_scenario(L, B, steps, unique=4)
    a = gpu.CSLAMBatch(B, L)
    a.set_state(sc.x0, sc.S0)
    xs_sync = []
    for s in range(steps):
        a.SLAM(sc.u[s], sc.z[s], sc.matched[s])
        xs_sync.append(a.get_x().copy())
    a.close()
    b = gpu.CSLAMBatch(B, L)
    b.set_state(sc.x0, sc.S0)
    hu = torch.from_numpy(sc.u).pin_memory()
    hz = torch.from_numpy(sc.z).pin_memory()
    hm = torch.from_numpy(sc.matched).pin_memory()
    hx = torch.empty((steps, B, 6 * L + 4), dtype=torch.float64).pin_memory()
    from cv_monoslam_b200 import capi
    for s in range(steps):
        capi.check(b._lib.srukf_step(b._h, hu[s].data_ptr(), hz[s].data_ptr(), hm[s].data_ptr()))
        b.get_x_async(hx[s].data_ptr())
    b.sync()
    for s in range(steps):
        assert np.array_equal(hx[s].numpy(), xs_sync[s]), s
    b.close()
