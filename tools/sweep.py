"""BASELINE config 4: state-dimension sweep on one GPU (run under gpurun).  Prints one JSON object per L."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from cv_monoslam_b200 import CSLAMBatch
import synth
from cv_monoslam_b200.slam import tri_pack

def run(L, B, steps=3, warmup=3):
    n = 6 * L + 4
    # The reference's independent per-feature downdates over-subtract the common robot information L times, so its
    # stability margin shrinks with L: at L = 100 the default synthetic odometry noise already pushes 1 filter in
    # ~8000 into a real GMW modification (the oracle agrees: max E = 3.7e-7 for that filter) and from there into the
    # chaotic regime, which the fused path hands to the slow reference-order fallback.  Throughput at large L is
    # therefore measured with a gentler control / odometry noise.
    noise = synth.Noise() if L < 80 else synth.Noise(control=(0.003, 0.001, 0.003), odo_sigma=(3e-4, 1.5e-4, 3e-4))
    sc = synth.make_scenario(L, B, steps + warmup, unique=4, dense_state=False, noise=noise)
    g = CSLAMBatch(B, L)
    xw = torch.from_numpy(sc.x0).cuda(); Sw = torch.from_numpy(tri_pack(sc.S0)).cuda()
    wof = torch.from_numpy(sc.meta["world_of"]).cuda()
    for b0 in range(0, B, 4096):
        nb = min(4096, B - b0); idx = wof[b0:b0 + nb]
        xs, Ss = xw[idx].contiguous(), Sw[idx].contiguous()
        torch.cuda.current_stream().synchronize()   # the handle's stream is not ordered after torch's gather
        g.set_state_dev(b0, nb, xs.data_ptr(), Ss.data_ptr())
    du = torch.from_numpy(sc.u).cuda(); dz = torch.from_numpy(sc.z).cuda(); dm = torch.from_numpy(sc.matched).cuda()
    st = torch.cuda.ExternalStream(g.stream())
    for s in range(warmup):
        g.SLAM_dev(du[s].data_ptr(), dz[s].data_ptr(), dm[s].data_ptr())
    g.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for s in range(warmup, warmup + steps):
        g.SLAM_dev(du[s].data_ptr(), dz[s].data_ptr(), dm[s].data_ptr())
    e1.record(st); g.sync(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    w = bench.flops_downdate(n, L) + bench.flops_gain(n, L) + bench.flops_predict(n, L)
    rate = B / (ms * 1e-3)
    fl = g.flags()
    out = dict(L=L, n=n, sigma_points=2 * (n + 5) + 1, filters=B, ms_per_step=ms, filter_steps_per_s=rate,
               mflop_per_step=w / 1e6, frac_fp64_peak=rate * w / (bench.FP64_PEAK_FALLBACK_TFLOPS * 1e12),
               flag_or=int(np.bitwise_or.reduce(fl)), n_fallback=int((fl & 32 != 0).sum()),
               finite=bool(np.isfinite(g.get_x()).all()))
    g.close()
    return out

if __name__ == "__main__":
    cfg = [(10, 262144), (20, 131072), (33, 65536), (50, 65536), (66, 32768), (100, 8192)]
    if len(sys.argv) > 1:
        cfg = [(int(a.split(":")[0]), int(a.split(":")[1])) for a in sys.argv[1:]]
    for L, B in cfg:
        print(json.dumps(run(L, B)), flush=True)
