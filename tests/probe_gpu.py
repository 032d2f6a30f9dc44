"""Developer probe (run under gpurun): GPU vs oracle on a small scenario, prints error metrics."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import oracle as O
from cv_monoslam_b200 import CSLAMBatch, capi
import synth

def relmax(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)

def run(L, B, steps, mode_gpu=0, split=False):
    sc = synth.make_scenario(L, B, steps, unique=min(B, 8))
    prm = capi.default_params(downdate_mode=mode_gpu)
    g = CSLAMBatch(B, L, prm)
    g.set_state(sc.x0, sc.S0)
    p = O.default_params(downdate_mode=1)
    x = sc.x0.copy(); S = sc.S0.copy()
    n = 6 * L + 4
    worst = (0, 0)
    for s in range(steps):
        if split:
            g.predictMotion(sc.u[s]); g.predictMeasurement(); g.KalmanUpdate(sc.z[s], sc.matched[s])
        else:
            g.SLAM(sc.u[s], sc.z[s], sc.matched[s])
        O.batch_step(p, x, S, sc.u[s:s+1], sc.z[s:s+1], sc.matched[s:s+1], 8)
        xg, Sg = g.get_state()
        Pg = np.einsum('bki,bkj->bij', Sg, Sg); Po = np.einsum('bki,bkj->bij', S, S)
        ex = max(relmax(xg[b], x[b]) for b in range(B)); eP = max(relmax(Pg[b], Po[b]) for b in range(B))
        worst = (max(worst[0], ex), max(worst[1], eP))
        if s < 3 or s == steps - 1:
            print(f"  L={L} step {s}: rel x {ex:.3e}  rel P {eP:.3e}  flags {np.bitwise_or.reduce(g.flags())}", flush=True)
    print(f"L={L} B={B} steps={steps} mode={mode_gpu} split={split}: worst rel x {worst[0]:.3e}, rel P {worst[1]:.3e}")
    g.close()

if __name__ == "__main__":
    run(3, 4, 5)
    run(8, 8, 10)
    run(8, 8, 5, mode_gpu=1)
    run(8, 8, 5, split=True)
    run(20, 8, 10)
