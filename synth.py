"""Synthetic ceiling-SLAM inputs for the batched SRUKF: priors, controls, measurements.

Host-side numpy only (input generation is not on the hot path).  The prior comes from a numpy
restatement of the reference's feature initialisation (MonoSLAM/SLAM.cpp:818-871, 1177-1334: an
unscented transform through the inverse-depth mapping, QR, permutation to canonical order) because
arbitrary well-conditioned priors make the reference's update diverge within a few frames
(SURVEY V1/V2).  The factor is then put in the form every post-update factor has in the reference,
S <- modifiedCholesky(S^T S) (SLAM.cpp:2152, 2197-2327), so it is unique up to rounding.

TEST / BENCH INPUT INFRASTRUCTURE, not part of the product package: bench.py, smoke() and tests/ import it for
inputs; tests also use its numpy restatements as an independent cross-check of the C oracle.  It is independent
of oracle/.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

SEED0 = 20140101


@dataclass
class Camera:
    # SLAM.cpp:329-337, image size assumed 640x480 (:312-313 reads it from the first frame)
    dx: float = 0.0028
    dy: float = 0.0028
    cx: float = 310.1129
    cy: float = 236.7526
    k1: float = 0.0001
    k2: float = 0.0
    f: float = 2.1735
    width: int = 640
    height: int = 480
    epsilon: float = 1e-13

    @property
    def f1(self):
        return self.f / self.dx

    @property
    def f2(self):
        return self.f / self.dy


@dataclass
class Noise:
    sigma_measure: float = 3.0                       # SLAM.cpp:189
    rho0: float = 1.0 / 3.0                          # :172-173
    sigma_rho: float = (1.0 / 3.0) / 2.0             # :190
    S4: tuple = (0.02, 0.02, 0.005, 0.02)            # :221-224
    odo_sigma: tuple = (1e-3, 5e-4, 1e-3)            # synthetic odometry noise
    pix_sigma: float = 1.0                           # synthetic measurement noise
    control: tuple = (0.005, 0.002, 0.005)           # true (rot1, trans, rot2) per step: 0.2 m-radius circle.
    # SURVEY 8(d) suggested (0.025, 0.005, 0.025); with the reference's a1..a4 = 8 that makes Mt ~ 1 cm per
    # frame and the reference's independent per-feature downdates over-subtract -> divergence in 3 frames
    # (measured with the oracle).  The smaller control keeps the reference filter stable.
    kp_radius: tuple = (30.0, 120.0)                 # key-point radius range [px] around the principal point
    ceiling: float = 3.0                             # m_deep, :172


def sample_weights(Na: int, weight_type: int = 0, alpha: float = 1e-3, beta: float = 2.0) -> dict:
    """calculateSampleParameter, SLAM.cpp:1050-1103."""
    if weight_type == 0:
        wm0 = 1.0 - Na / 3.0
        wc0 = 1.0 - Na / 3.0
        wi = (1.0 - wc0) / (2 * Na)
        gamma = np.sqrt(Na / (1.0 - wm0))
    elif weight_type == 1:
        lam = alpha ** 2 * (Na + 0.0) - Na
        gamma = np.sqrt(Na + lam)
        wm0 = lam / (Na + lam)
        wc0 = wm0 + (1 - alpha ** 2 + beta)
        wi = 1.0 / (2 * (Na + lam))
    else:
        gamma = np.sqrt(3.0 * Na / 2.0)
        wm0 = wc0 = 1.0 / 3.0
        wi = 1.0 / (3.0 * Na)
    return dict(gamma=float(gamma), wm0=float(wm0), wc0=float(wc0), wi=float(wi), wi_sr=float(np.sqrt(abs(wi))))


# ------------------------------------------------------------------------------------------------
# camera chain (vectorised)
# ------------------------------------------------------------------------------------------------
def undistort(cam: Camera, uvd):
    """undistortOnePointRW, SLAM.cpp:3224-3236."""
    uvd = np.asarray(uvd, dtype=np.float64)
    xd = (uvd[..., 0] - cam.cx) * cam.dx
    yd = (uvd[..., 1] - cam.cy) * cam.dy
    rd = np.sqrt(xd * xd + yd * yd)
    d = 1 + cam.k1 * rd ** 2 + cam.k2 * rd ** 4
    return np.stack([cam.cx + xd * d / cam.dx, cam.cy + yd * d / cam.dy], axis=-1)


def distort(cam: Camera, uvu, iters: int = 100):
    """distortOnePointRW, SLAM.cpp:3177-3213 (fixed-iteration Newton; stops early when stationary)."""
    uvu = np.asarray(uvu, dtype=np.float64)
    xu = (uvu[..., 0] - cam.cx) * cam.dx
    yu = (uvu[..., 1] - cam.cy) * cam.dy
    ru = np.sqrt(xu * xu + yu * yu)
    rd = ru / (1 + cam.k1 * ru * ru + cam.k2 * ru ** 4)
    for _ in range(iters):
        f = rd + cam.k1 * rd ** 3 + cam.k2 * rd ** 5 - ru
        ff = 1.0 + 3.0 * cam.k1 * rd * rd + 5.0 * cam.k2 * rd ** 4
        nrd = rd - f / ff
        if np.array_equal(nrd, rd):
            break
        rd = nrd
    d = 1 + cam.k1 * rd * rd + cam.k2 * rd ** 4
    d = np.where(d == 0, cam.epsilon, d)
    ox = cam.cx + xu / d / cam.dx
    oy = cam.cy + yu / d / cam.dy
    vis = (ox >= 0) & (ox <= cam.width) & (oy >= 0) & (oy <= cam.height)
    return np.stack([np.where(vis, ox, 0.0), np.where(vis, oy, 0.0)], axis=-1)


def _camera_to_pixel(cam: Camera, Hr, err=None):
    """coordinatesCamera2Image, SLAM.cpp:3324-3347 (x/y swap is the reference's), then distortion."""
    e0 = 0.0 if err is None else err[..., 0]
    e1 = 0.0 if err is None else err[..., 1]
    z = Hr[..., 2]
    zs = np.where(z == 0, 1.0, z)
    uy = cam.cx + cam.f1 * Hr[..., 0] / zs + e0
    ux = cam.cy + cam.f2 * Hr[..., 1] / zs + e1
    bad = (z == 0) | (ux < 10) | (ux > cam.width - 10) | (uy < 10) | (uy > cam.height - 10)
    uvu = np.stack([np.where(bad, 0.0, ux), np.where(bad, 0.0, uy)], axis=-1)
    return distort(cam, uvu)


def _world_to_camera(Hw, theta):
    """coordinatesWorld2Camera with Rcw = Rwc^-1 (SLAM.cpp:1642-1643, 3289-3292)."""
    c, s = np.cos(theta), np.sin(theta)
    return np.stack([c * Hw[..., 0] + s * Hw[..., 1], -s * Hw[..., 0] + c * Hw[..., 1], Hw[..., 2]], axis=-1)


def project_state(cam: Camera, feat, pos, theta, err=None):
    """Inverse-depth feature (…,6) seen from robot position (…,3) and heading: State2World
    (SLAM.cpp:3250-3276) -> World2Camera -> Camera2Image -> distortion."""
    xi, yi, zi, th, ph, rho = (feat[..., k] for k in range(6))
    Hw = np.stack([xi + 1 / rho * np.cos(ph) * np.sin(th) - pos[..., 0],
                   yi - 1 / rho * np.sin(ph) - pos[..., 1],
                   zi + 1 / rho * np.cos(ph) * np.cos(th) - pos[..., 2]], axis=-1)
    return _camera_to_pixel(cam, _world_to_camera(Hw, theta), err)


def project_world(cam: Camera, Pw, pos, theta):
    """Cartesian world point(s) seen from a robot pose (used to synthesise measurements)."""
    return _camera_to_pixel(cam, _world_to_camera(Pw - pos, theta))


def backproject_direction(cam: Camera, uvd, theta):
    """undistort -> coordinatesImage2Camera (SLAM.cpp:3358-3363) -> Camera2World (:3382-3387)."""
    uvu = undistort(cam, uvd)
    hx = (uvu[..., 1] - cam.cx) / cam.f1
    hy = (uvu[..., 0] - cam.cy) / cam.f2
    c, s = np.cos(theta), np.sin(theta)
    return np.stack([c * hx - s * hy, s * hx + c * hy, np.ones_like(hx)], axis=-1)


# ------------------------------------------------------------------------------------------------
# modified Cholesky (numpy) and feature initialisation
# ------------------------------------------------------------------------------------------------
def mchol(G: np.ndarray, epsilon: float = 1e-13):
    """Gill-Murray-Wright modified Cholesky as in SLAM.cpp:2197-2327.  Returns (S, E)."""
    G = np.asarray(G, dtype=np.float64)
    n = G.shape[0]
    gamma = np.max(np.diag(G))
    off = G - np.diag(np.diag(G))
    zi = np.max(off)
    nu = max(1.0, np.sqrt(n * n - 1.0))
    beta2 = max(gamma, zi / nu, 1e-15)
    C = np.tril(G).copy()
    S = np.zeros((n, n))
    E = np.zeros(n)
    for j in range(n):
        col = C[j:, j]
        theta = np.max(np.abs(col[1:])) if j < n - 1 else 0.0
        d = max(epsilon, abs(col[0]), theta * theta / beta2)
        E[j] = d - col[0]
        sd = np.sqrt(d)
        S[j, j] = sd
        if j < n - 1:
            l = col[1:] / d
            S[j, j + 1:] = sd * l
            C[j + 1:, j + 1:] -= np.tril(np.outer(col[1:], l))
    return S, E


def init_prior(cam: Camera, noise: Noise, x4, kp, weight_type: int = 0, canonical: bool = True):
    """Feature initialisation at frame 1 (SLAM.cpp:818-871, 1177-1334) for M key-points kp [M,2].

    Returns x [6M+4], S [6M+4, 6M+4] (upper triangular, canonical state order)."""
    kp = np.asarray(kp, dtype=np.float64).reshape(-1, 2)
    M = kp.shape[0]
    dim = 4
    Na = dim + 3 * M
    w = sample_weights(Na, weight_type)
    mu = np.concatenate([np.asarray(x4, dtype=np.float64),
                         np.column_stack([kp, np.full(M, noise.rho0)]).ravel()])
    sr = np.zeros((Na, Na))
    sr[:4, :4] = np.diag(noise.S4)
    for i in range(M):
        sr[4 + 3 * i, 4 + 3 * i] = noise.sigma_measure
        sr[5 + 3 * i, 5 + 3 * i] = noise.sigma_measure
        sr[6 + 3 * i, 6 + 3 * i] = noise.sigma_rho
    # sigma points as columns (:1148-1162)
    sig = np.empty((Na, 2 * Na + 1))
    sig[:, 0] = mu
    sig[:, 1:Na + 1] = mu[:, None] + w["gamma"] * sr.T
    sig[:, Na + 1:] = mu[:, None] - w["gamma"] * sr.T
    # mapping (:1201-1242)
    theta = sig[3]                                            # [P]
    pix = sig[4:].reshape(M, 3, -1)                           # [M,3,P]
    uvd = np.stack([pix[:, 0], pix[:, 1]], axis=-1)           # [M,P,2]
    Hw = backproject_direction(cam, uvd, theta[None, :])      # [M,P,3]
    ang = np.stack([np.arctan2(Hw[..., 0], Hw[..., 2]),
                    np.arctan2(-Hw[..., 1], np.sqrt(Hw[..., 0] ** 2 + Hw[..., 2] ** 2)),
                    pix[:, 2]], axis=1)                       # [M,3,P]
    P = 2 * Na + 1
    out = np.empty((dim + 6 * M, P))
    out[:dim] = sig[:dim]
    out[dim:dim + 3 * M] = ang.reshape(3 * M, P)
    out[dim + 3 * M:] = np.tile(sig[0:3], (M, 1))
    wv = np.full(P, w["wi"])
    wv[0] = w["wm0"]
    mu_angle = ang.reshape(3 * M, P) @ wv
    xdis = np.concatenate([mu[:4], mu_angle, np.tile(mu[0:3], M)])
    A = w["wi_sr"] * (out[:, 1:] - out[:, :1]).T              # [2Na, dimNew]
    R = np.linalg.qr(A, mode="r")
    # permutation to canonical order (:1303-1334)
    dimNew = dim + 6 * M
    src = np.empty(dimNew, dtype=int)
    src[dimNew - 4:] = np.arange(4)
    for i in range(M):
        src[6 * i:6 * i + 3] = 4 + 3 * M + 3 * i + np.arange(3)
        src[6 * i + 3:6 * i + 6] = 4 + 3 * i + np.arange(3)
    x = xdis[src]
    Pcov = (R.T @ R)[np.ix_(src, src)]
    if canonical:
        S, _ = mchol(Pcov, cam.epsilon)
    else:
        S = np.linalg.qr(R[:, src], mode="r")
    return x, S


# ------------------------------------------------------------------------------------------------
# scenarios
# ------------------------------------------------------------------------------------------------
@dataclass
class Scenario:
    L: int
    B: int
    steps: int
    x0: np.ndarray          # [B, n]
    S0: np.ndarray          # [B, n, n]
    u: np.ndarray           # [steps, B, 3]      odometry controls (rot1, trans, rot2)
    z: np.ndarray           # [steps, B, L, 2]   matched pixels
    matched: np.ndarray     # [steps, B, L] uint8
    truth: np.ndarray       # [steps, B, 3]      true (x, y, theta) after each step
    meta: dict = field(default_factory=dict)


def make_world(L: int, world_id: int, steps: int, seed0: int, cam: Camera, noise: Noise):
    """One synthetic world: prior (x0, S0), true trajectory and noise-free measurements."""
    rng = np.random.default_rng(seed0 + world_id)
    theta0 = rng.uniform(-np.pi, np.pi)
    x4 = np.array([0.0, 0.0, 0.0, theta0])
    pose = x4 + np.asarray(noise.S4) * rng.standard_normal(4)              # true initial pose
    r = rng.uniform(noise.kp_radius[0], noise.kp_radius[1], L)
    a = rng.uniform(0.0, 2 * np.pi, L)
    kp = np.stack([cam.cx + r * np.cos(a), cam.cy + r * np.sin(a)], axis=-1)
    # true landmarks: back-project the noise-free key-points from the true pose onto the ceiling
    kp_true = kp - noise.pix_sigma * rng.standard_normal((L, 2))
    d = backproject_direction(cam, kp_true, pose[3])
    t = (noise.ceiling - pose[2]) / d[:, 2]
    Pw = pose[None, 0:3] + t[:, None] * d
    x0, S0 = init_prior(cam, noise, x4, kp)
    truth = np.empty((steps, 3))
    zc = np.empty((steps, L, 2))
    r1, tr, r2 = noise.control
    for s in range(steps):                                                  # odometry motion model, :1518-1521
        pose[0] += tr * np.cos(pose[3] + r1)
        pose[1] += tr * np.sin(pose[3] + r1)
        pose[3] += r1 + r2
        truth[s] = (pose[0], pose[1], pose[3])
        zc[s] = project_world(cam, Pw, pose[None, 0:3], pose[3])
    return x0, S0, truth, zc


def make_scenario(L: int, B: int, steps: int, unique: int | None = None, seed0: int = SEED0,
                  cam: Camera | None = None, noise: Noise | None = None, match_prob: float = 1.0,
                  first_filter: int = 0, dense_state: bool = True) -> Scenario:
    """B filters (global ids first_filter .. first_filter+B-1), `steps` frames.

    Global filter g lives in world g % unique (priors/worlds are generated once per world and replicated:
    SURVEY 8(d) allows throughput runs to replicate one trajectory's inputs) and draws its odometry and
    pixel noise from seed seed0 + 1000003*(g+1), so its inputs do not depend on how the batch is sharded."""
    cam = cam or Camera()
    noise = noise or Noise()
    n = 6 * L + 4
    U = (first_filter + B) if unique is None else max(1, unique)
    ids = np.arange(first_filter, first_filter + B)
    worlds = {w: make_world(L, w, steps, seed0, cam, noise) for w in sorted(set(int(g % U) for g in ids))}
    wkeys = sorted(worlds)
    if dense_state:
        x0 = np.empty((B, n))
        S0 = np.empty((B, n, n))
    else:  # per-world priors only (x0 [U',n], S0 [U',n,n]) plus meta["world_of"]: large batches replicate on device
        x0 = np.stack([worlds[w][0] for w in wkeys])
        S0 = np.stack([worlds[w][1] for w in wkeys])
    world_of = np.array([wkeys.index(int(g % U)) for g in ids], dtype=np.int64)
    u = np.empty((steps, B, 3))
    z = np.empty((steps, B, L, 2))
    truth = np.empty((steps, B, 3))
    matched = np.ones((steps, B, L), dtype=np.uint8)
    ctrl = np.asarray(noise.control)
    odo = np.asarray(noise.odo_sigma)
    for b, g in enumerate(ids):
        wx0, wS0, wtruth, wz = worlds[int(g % U)]
        truth[:, b] = wtruth
        if dense_state:
            x0[b], S0[b] = wx0, wS0
        rng = np.random.default_rng(seed0 + 1_000_003 * (int(g) + 1))
        u[:, b] = ctrl + odo * rng.standard_normal((steps, 3))
        z[:, b] = wz + noise.pix_sigma * rng.standard_normal((steps, L, 2))
        if match_prob < 1.0:
            matched[:, b] = (rng.uniform(size=(steps, L)) < match_prob).astype(np.uint8)
        # a landmark whose true pixel is outside the view (project_world returns (0, 0), as the reference's
        # Camera2Image / distortion zeroing does) cannot be matched by dataAssociation (SLAM.cpp:1946-2001)
        matched[:, b] &= ((wz[..., 0] != 0.0) & (wz[..., 1] != 0.0)).astype(np.uint8)
    return Scenario(L=L, B=B, steps=steps, x0=x0, S0=S0, u=u, z=z, matched=matched, truth=truth,
                    meta=dict(unique=U, seed0=seed0, match_prob=match_prob, first_filter=first_filter,
                              world_of=world_of, dense_state=dense_state))
