"""Independent numpy restatement of one literal SRUKF frame (MonoSLAM/SLAM.cpp:1430-1775, 2020-2327).

Second, independent implementation used to cross-check the C oracle (two restatements must
agree <= 1e-12).  Deliberately different machinery: dense sigma matrices in numpy, LAPACK QR
(np.linalg.qr) instead of the restated GSL Householder, synth.mchol / synth.project_state instead of
the C camera chain.  TEST INFRASTRUCTURE ONLY.
"""
import numpy as np

import synth


def step(x, S, u, z, matched, cam=None, noise_sigma=3.0, a=(8, 8, 8, 8), weight_type=0, eps=1e-13):
    cam = cam or synth.Camera()
    x = np.array(x, dtype=np.float64)
    S = np.array(S, dtype=np.float64)
    n = x.size
    L = (n - 4) // 6
    Na = n + 5
    w = synth.sample_weights(Na, weight_type)
    g, wm0, wc0, wi = w["gamma"], w["wm0"], w["wc0"], w["wi"]
    r1, tr, r2 = u
    Mt = np.diag([a[0] * r1 * r1 + a[1] * tr * tr, a[2] * tr * tr + a[3] * r1 * r1 + a[3] * r2 * r2,
                  a[0] * r2 * r2 + a[1] * tr * tr])
    sr = np.zeros((Na, Na))
    sr[:n, :n] = S
    sr[n:n + 3, n:n + 3] = Mt
    sr[n + 3:, n + 3:] = np.eye(2) * noise_sigma
    mu = np.concatenate([x, np.zeros(5)])
    sig = np.empty((Na, 2 * Na + 1))
    sig[:, 0] = mu
    sig[:, 1:Na + 1] = mu[:, None] + g * sr.T
    sig[:, Na + 1:] = mu[:, None] - g * sr.T
    # motion (:1476-1532)
    rot1 = r1 - sig[n]
    trans = tr - sig[n + 1]
    rot2 = r2 - sig[n + 2]
    th = sig[n - 1].copy()
    sig[n - 4] += trans * np.cos(th + rot1)
    sig[n - 3] += trans * np.sin(th + rot1)
    sig[n - 1] += rot1 + rot2
    wv = np.full(2 * Na + 1, wi)
    wv[0] = wm0
    x[n - 4:] = sig[n - 4:n] @ wv
    A = np.sqrt(wi) * (sig[:n, 1:] - sig[:n, :1]).T
    S = np.linalg.qr(A, mode="r")
    # measurement (:1615-1682)
    P = 2 * Na + 1
    feats = sig[:6 * L].reshape(L, 6, P).transpose(0, 2, 1)          # [L,P,6]
    pos = sig[n - 4:n - 1].T[None]                                  # [1,P,3]
    err = sig[n + 3:n + 5].T[None]                                  # [1,P,2]
    pix = synth.project_state(cam, feats, pos, sig[n - 1][None], err)  # [L,P,2]
    hbar = np.einsum("lpc,p->lc", pix, wv)
    si = []
    for j in range(L):
        Aj = np.sqrt(wi) * (pix[j, 1:] - pix[j, :1])
        si.append(np.linalg.qr(Aj, mode="r"))
    # update (:2048-2096)
    wc = wv.copy()
    wc[0] = wc0
    for j in range(L):
        if not matched[j] or hbar[j, 0] == 0 or hbar[j, 1] == 0:
            continue
        d1 = sig[:n] - x[:, None]
        d2 = (pix[j] - hbar[j]).T
        Pxy = (d1 * wc) @ d2.T
        sii = np.linalg.inv(si[j])
        Ki = Pxy @ sii @ sii.T
        x = x + Ki @ (z[j] - hbar[j])
        U = Ki @ si[j].T
        for c in range(2):
            G = S.T @ S - np.outer(U[:, c], U[:, c])
            S, _ = synth.mchol(G, eps)
    return x, S, hbar
