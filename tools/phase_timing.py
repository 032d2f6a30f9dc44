"""Diagnostics: per-phase SM cycles of k_update (run under gpurun with SRUKF_PHASE_TIMING=1)."""
import os, sys
os.environ["SRUKF_PHASE_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cv_monoslam_b200 import CSLAMBatch, capi
import synth
L = int(sys.argv[1]) if len(sys.argv) > 1 else 50
B = int(sys.argv[2]) if len(sys.argv) > 2 else 592
sc = synth.make_scenario(L, B, 3, unique=2)
g = CSLAMBatch(B, L)
g.set_state(sc.x0, sc.S0)
for s in range(3):
    g.SLAM(sc.u[s], sc.z[s], sc.matched[s])
g.sync()
out = np.zeros(16, dtype=np.uint64)
capi.check(g._lib.srukf_get_phase_cycles(g._h, capi.ptr(out)))
n = float(out[7])
names = ["K loop", "post-K barrier", "panel store+sync", "factor panel", "write S_new", "end barrier"]
tot = float(out[:6].sum())
for nm, v in zip(names, out[:6]):
    print(f"{nm:18s} {float(v)/n:12.0f} cycles/filter  {100*float(v)/tot:5.1f}%")
print(f"total {tot/n:.0f} cycles per filter-update ({int(n)} CTAs)")
kn = ["acquire (empty wait)", "copy issue", "data wait (full)", "DMMA", "release"]
nch = float(out[13]) / n
for nm, v in zip(kn, out[8:13]):
    print(f"  K loop / {nm:22s} {float(v)/n:12.0f} cycles/filter   {float(v)/max(float(out[13]),1):8.0f} per chunk")
print(f"  chunks per filter {nch:.0f}")
