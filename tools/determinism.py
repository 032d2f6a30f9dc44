"""Diagnostics: run the same scenario twice and compare the final state bit for bit (run under gpurun)."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cv_monoslam_b200 import CSLAMBatch
import synth

def run(L, B, steps):
    noise = synth.Noise() if L < 80 else synth.Noise(control=(0.003, 0.001, 0.003), odo_sigma=(3e-4, 1.5e-4, 3e-4))
    sc = synth.make_scenario(L, B, steps, unique=4, noise=noise)
    out = []
    for rep in range(2):
        g = CSLAMBatch(B, L)
        g.set_state(sc.x0, sc.S0)
        for s in range(steps):
            g.SLAM(sc.u[s], sc.z[s], sc.matched[s])
        x, S = g.get_state(dense=False)
        fl = g.flags()
        out.append((x.copy(), S.copy(), fl.copy()))
        g.close()
    (x0, S0, f0), (x1, S1, f1) = out
    dx = np.abs(x0 - x1).max(); dS = np.abs(S0 - S1).max()
    print(f"L={L} B={B} steps={steps}: max|dx|={dx:.3e} max|dS|={dS:.3e} flags differ={(f0 != f1).sum()} "
          f"fallbacks={(f0 & 32 != 0).sum()}/{(f1 & 32 != 0).sum()} bitwise_equal={np.array_equal(x0, x1) and np.array_equal(S0, S1)}")
    if dS > 0:
        b = np.argwhere(np.abs(S0 - S1).reshape(B, -1).max(axis=1) > 0).ravel()
        print("  filters that differ:", b[:10], "count", len(b))

if __name__ == "__main__":
    for a in sys.argv[1:]:
        L, B, steps = (int(v) for v in a.split(":"))
        run(L, B, steps)
