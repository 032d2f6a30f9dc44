#!/usr/bin/env python
"""Quick GPU-vs-oracle parity probe over a few state sizes (development aid; the real tests are tests/test_gpu_parity.py).
usage: python tools/quick_parity.py [L:B:steps ...]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O
import synth
from cv_monoslam_b200 import CSLAMBatch, capi

def cov(S): return np.einsum("...ki,...kj->...ij", S, S)
cases = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(3, 3, 3), (5, 4, 3), (8, 4, 4), (20, 4, 4), (50, 2, 2)]
ok = True
for (L, B, steps) in cases:
    sc = synth.make_scenario(L, B, steps, unique=B)
    g = CSLAMBatch(B, L)
    g.set_state(sc.x0, sc.S0)
    p = O.default_params(downdate_mode=1)
    x, S = sc.x0.copy(), sc.S0.copy()
    t0 = time.time()
    for s in range(steps):
        g.SLAM(sc.u[s], sc.z[s], sc.matched[s])
        O.batch_step(p, x, S, sc.u[s:s + 1], sc.z[s:s + 1], sc.matched[s:s + 1], os.cpu_count() or 1)
        xg, Sg = g.get_state()
        ex = max(np.abs(xg[b] - x[b]).max() / np.abs(x[b]).max() for b in range(B))
        eP = max(np.abs(cov(Sg[b]) - cov(S[b])).max() / np.abs(cov(S[b])).max() for b in range(B))
        print(f"L={L} B={B} step {s}: rel err x {ex:.2e} P {eP:.2e} flags {np.bitwise_or.reduce(g.flags()):#x}", flush=True)
        ok &= bool(ex < 1e-9 and eP < 1e-9)
    g.close()
print("QUICK PARITY", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
