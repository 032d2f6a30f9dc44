#!/usr/bin/env python
"""bench.py -- SRUKF filter-steps/s (predictMotion + predictMeasurement + KalmanUpdate, FP64).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

One "step" = one frame of the hot path for every filter of the batch.  Workload at N GPUs: BASELINE.json
configs[2] per GPU (65,536 filters x 50 landmarks, n = 304, 619 sigma points), i.e. weak scaling towards
configs[4] (524,288 filters at N = 8).  Synthetic inputs: synth.py.

Prints ONE JSON line (rank 0).  `value` is timed with CUDA events on the library's stream with inputs
resident in HBM; `e2e` goes through the host-pointer C ABI (pinned host buffers, H2D of the step's inputs
and D2H of m_X_k inside the timed region, overlapped with the neighbouring frames by the library's copy streams).  The roofline entry is for the dominant kernel (k_update),
timed live with CUDA events by the library (srukf_set_profiling).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FP64_PEAK_FALLBACK_TFLOPS = 37.2   # only if the in-run measurement fails: profiles/r01_fp64_peak.json (DMMA m8n8k4)


def measured_peaks() -> dict:
    """MEASURED_PEAKS.json (driver-written): HBM copy bandwidth of this pool's B200s.  It has no FP64 entry: the FP64
    tensor-pipe peak is measured by the library in this very process (srukf_fp64_peak) next to the timed region."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


def committed_traffic(kernel: str, L: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per FILTER of `kernel` from the latest committed `ncu --set full`
    summary of this command (profiles/*_traffic.json, written by tools/ncu_summary.py --json); None if there is none
    for this L.  Never a constant in this file."""
    import glob
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json"))):
        try:
            with open(path) as f:
                rec = json.load(f)
        except (OSError, ValueError):
            continue
        k = rec.get("kernels", {}).get(kernel)
        if k and int(rec.get("landmarks", -1)) == L:
            best = (k["dram_bytes_per_filter"], os.path.relpath(path, ROOT))
    return best


# ------------------------------------------------------------------------------------------------------
# work model (DESIGN.md "Work per filter-step"): flops of the formulation that is actually executed
# ------------------------------------------------------------------------------------------------------
def flops_downdate(n: int, L: int) -> float:
    """k_update: multiply-adds of G = P - U U^T on the lower triangle (P = S^T S is carried between steps, so the
    reference's dense S^T S product is NOT counted) plus the modified Cholesky of G.  The fused left-looking kernel
    performs these products in panel order; padding and masked tiles are not counted."""
    form = n * (n + 1) / 2.0 * (2 * L) * 2.0
    mchol = sum((n - j - 1) * (n - j) / 2.0 * 2.0 + 3.0 * (n - j) for j in range(n))
    return form + mchol


def flops_gain(n: int, L: int) -> float:
    nf = n - 4
    return nf * (nf + 1) / 2.0 * 2 * L * 2.0 + 12.0 * n * L + 8.0 * nf * L


def flops_predict(n: int, L: int) -> float:
    Na = n + 5
    P = 2 * Na + 1
    nf = n - 4
    # projections (~100 flop each incl. sincos) + sums + robot rows of the carried covariance (S_ff^T E_f)
    return 100.0 * P * L + 40.0 * P + 26.0 * 2 * Na * L + nf * (nf + 1) / 2.0 * 4 * 2.0


def algorithmic_bytes(n: int, L: int) -> float:
    return 8.0 * (n * (n + 1) + 2 * n + 3 + 2 * L)          # packed S in+out, x in+out, u, z


# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        for line in self.f.read().splitlines():
            c = [t.strip() for t in line.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1])); pw.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def workload_string(B: int, L: int) -> str:
    n = 6 * L + 4
    return (f"batched SRUKF predict+update, {B} filters x {L} landmarks per GPU (n={n}, {2 * (n + 5) + 1} sigma points), "
            "FP64 (BASELINE configs[2])")


def config_dict(B: int, L: int, unique: int, world: int, downdate_mode: int) -> dict:
    """the `config` object of BOTH arms (the reference arm times a sample of this workload, see cpu_baseline.sample)"""
    n = 6 * L + 4
    ntri = n * (n + 1) // 2
    return {"workload": workload_string(B, L),
            "filters_per_gpu": B, "landmarks": L, "state_dim": n, "sigma_points": 2 * (n + 5) + 1,
            "distinct_worlds": int(unique), "parallelism": f"filters sharded over {world} GPU(s)",
            "l2": f"inputs larger than L2: {B * ntri * 8 / 2**30:.1f} GiB of packed S per GPU streamed per step",
            "downdate_mode": downdate_mode,
            **({"forced_fallback_ppm": int(os.environ["SRUKF_FORCE_FALLBACK_PPM"])}
               if os.environ.get("SRUKF_FORCE_FALLBACK_PPM") else {})}


def cpu_reference_run(L: int, filters: int, steps: int, warmup: int, threads: int, seed_first: int = 0):
    """Times the reference's CPU implementation of the path on `threads` host threads, one filter per thread at a time:
    predictMotion + predictMeasurement + KalmanUpdate with materialised sigma matrices, the Householder QR of the
    2Na x n matrix, and one dense S^T S + modified Cholesky per U column.

    kind "reference": oracle/_ref/libsrukf_ref.so -- the bodies of the reference's own functions (extracted verbatim from
    MonoSLAM/SLAM.cpp where /root/reference exists; the prebuilt library travels to the GPU box) over the cv::Mat stand-in
    of oracle/ref_shim.  kind "port": the C restatement oracle/srukf_oracle.c (bit-identical results, ~4x faster: no
    temporaries), used when the reference library is absent.  Returns (kind, seconds per step list)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import synth
    sc = synth.make_scenario(L, filters, warmup + steps, unique=min(filters, 4), first_filter=seed_first)
    times = []
    import ref as R  # oracle/ref.py: the one other place bench.py executes oracle/ -- as the measured CPU baseline
    if R.available():
        import threading
        R.lib()
        slams = []
        for b in range(filters):
            r = R.Slam()
            r.set_state(sc.x0[b], sc.S0[b])
            slams.append(r)
        odo = [[R.control_to_odometry(sc.u[s, b]) for b in range(filters)] for s in range(warmup + steps)]

        def work(tid, s):
            for b in range(tid, filters, threads):
                r = slams[b]
                r.predict_motion_odometry(*odo[s][b])
                r.predict_measurement()
                r.kalman_update(sc.z[s, b], sc.matched[s, b])

        for s in range(warmup + steps):
            t0 = time.perf_counter()
            ths = [threading.Thread(target=work, args=(t, s)) for t in range(threads)]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
            dt = time.perf_counter() - t0
            if s >= warmup:
                times.append(dt)
        x, _ = slams[0].get_state()
        assert np.isfinite(x).all()
        return "reference", times
    import oracle as O
    O.build()
    p = O.default_params(downdate_mode=2)
    x, S = sc.x0.copy(), sc.S0.copy()
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        O.batch_step(p, x, S, sc.u[s:s + 1], sc.z[s:s + 1], sc.matched[s:s + 1], threads)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    assert np.isfinite(x).all()
    return "port", times


CPU_KIND_TEXT = {
    "reference": "the reference's own SLAM.cpp function bodies (oracle/_ref, g++ -O2, cv::Mat stand-in of oracle/ref_shim)",
    "port": "oracle literal mode (C restatement of SLAM.cpp: dense S^T S + GMW per U column), gcc -O2",
}


def run_reference(args, ctx, out):
    """--impl reference: the reference's CPU implementation with all host threads; each step is a bounded SAMPLE of the
    workload (one filter per host thread), the metric is filter-steps/s of that sample."""
    if ctx.rank != 0:
        return
    cores = os.cpu_count() or 1
    L = args.landmarks
    filters = cores
    kind, times = cpu_reference_run(L, filters, args.steps, args.warmup, cores)
    tot = float(sum(times))
    value = filters * len(times) / tot
    n = 6 * L + 4
    line = {
        "impl": "reference", "metric": "srukf_filter_steps_per_sec", "value": value, "unit": "filter-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.filters, L, min(args.unique, args.filters), args.gpus, args.downdate_mode),
        "cpu_baseline": {"value": value, "unit": "filter-steps/s", "cores": cores, "kind": kind,
                         "sample": f"{filters} filters (one per host thread) x {len(times)} steps of the same workload: "
                                   + CPU_KIND_TEXT[kind]},
        "e2e": {"value": value, "unit": "filter-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    out.emit(json.dumps(line))


def load_priors(g, sc, B):
    """upload the distinct worlds' priors and replicate them over the batch on the device"""
    import torch
    from cv_monoslam_b200.slam import tri_pack
    xw = torch.from_numpy(sc.x0).cuda()
    Sw = torch.from_numpy(tri_pack(sc.S0)).cuda()
    wof = torch.from_numpy(sc.meta["world_of"]).cuda()
    slab = 4096
    for b0 in range(0, B, slab):
        nb = min(slab, B - b0)
        idx = wof[b0:b0 + nb]
        xs, Ss = xw[idx].contiguous(), Sw[idx].contiguous()
        torch.cuda.current_stream().synchronize()   # the handle's stream is not ordered after torch's gather
        g.set_state_dev(b0, nb, xs.data_ptr(), Ss.data_ptr())


def sweep_point(L, B, steps, warmup, peak, parity_max_L, dev, gentle=False):
    """BASELINE config 4, one state dimension: device-resident throughput of B filters + a parity spot check."""
    import torch
    from cv_monoslam_b200 import CSLAMBatch, capi
    import synth
    n = 6 * L + 4
    # The reference's independent per-feature downdates re-subtract the common robot information L times, so its
    # stability margin shrinks with L (profiles/r01_sweep_config4.md); `gentle` scales the synthetic control and odometry
    # noise down for the very large maps so that the run measures the fused path, not the reference-order fallback.
    noise = synth.Noise(control=(0.003, 0.001, 0.003), odo_sigma=(3e-4, 1.5e-4, 3e-4)) if gentle else synth.Noise()
    sc = synth.make_scenario(L, B, steps + warmup, unique=4, dense_state=False, noise=noise)
    g = CSLAMBatch(B, L, device=dev)
    load_priors(g, sc, B)
    du, dz, dm = (torch.from_numpy(a).cuda() for a in (sc.u, sc.z, sc.matched))
    torch.cuda.synchronize()
    st = torch.cuda.ExternalStream(g.stream(), device=dev)
    for s in range(warmup):
        g.SLAM_dev(du[s].data_ptr(), dz[s].data_ptr(), dm[s].data_ptr())
    g.sync()
    g.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for s in range(warmup, warmup + steps):
        g.SLAM_dev(du[s].data_ptr(), dz[s].data_ptr(), dm[s].data_ptr())
    g.sync()
    e1.record(st)
    g.sync()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    kms, _ = g.kernel_times()
    g.set_profiling(False)
    fl = g.flags()
    finite = bool(np.isfinite(g.get_x()).all())
    g.close()
    del du, dz, dm
    w = flops_downdate(n, L) + flops_gain(n, L) + flops_predict(n, L)
    rate = B / (ms * 1e-3)
    # ---- parity spot check on the same worlds: 2 filters x 2 frames against the CPU oracle (up to parity_max_L; the
    #      oracle's reference-order update costs 2L dense factorisations per frame), beyond that against the library's own
    #      reference-arithmetic path (downdate_mode 2: unblocked DFMA modified Cholesky in global memory)
    nchk, fr = 2, 2
    scp = synth.make_scenario(L, nchk, fr, unique=nchk, noise=noise)
    gp = CSLAMBatch(nchk, L, device=dev)
    gp.set_state(scp.x0, scp.S0)
    for s in range(fr):
        gp.SLAM(scp.u[s], scp.z[s], scp.matched[s])
    xg, Sg = gp.get_state()
    gp.close()
    if L <= parity_max_L:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as O  # checker of the sweep's spot check, not the thing measured
        x, S = scp.x0.copy(), scp.S0.copy()
        O.batch_step(O.default_params(downdate_mode=1), x, S, scp.u, scp.z, scp.matched, nchk)
        against = "cpu oracle (reference order)"
    else:
        g2 = CSLAMBatch(nchk, L, capi.default_params(downdate_mode=2), device=dev)
        g2.set_state(scp.x0, scp.S0)
        for s in range(fr):
            g2.SLAM(scp.u[s], scp.z[s], scp.matched[s])
        x, S = g2.get_state()
        g2.close()
        against = "library downdate_mode 2 (unblocked modified Cholesky, DFMA)"
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))   # noqa: E731
    ex = max(rel(xg[i], x[i]) for i in range(nchk))
    eP = max(rel(Sg[i].T @ Sg[i], S[i].T @ S[i]) for i in range(nchk))
    return {"metric": "srukf_filter_steps_per_sec", "value": rate, "unit": "filter-steps/s", "landmarks": L, "state_dim": n,
            "sigma_points": 2 * (n + 5) + 1, "filters": B, "steps": steps, "warmup": warmup, "ms_per_step": ms,
            "noise": "gentle (control and odometry noise / 8)" if gentle else "default",
            "kernel_ms_per_step": {"k_predict": float(kms[0]) / steps, "k_gain": float(kms[1]) / steps,
                                   "k_update": float(kms[2]) / steps},
            "mflop_per_filter_step": w / 1e6, "frac_fp64_peak": rate * w / (peak * 1e12), "peak_tflops": peak,
            "flag_or": int(np.bitwise_or.reduce(fl)), "n_fallback": int((fl & 32 != 0).sum()), "finite": finite,
            "parity": {"filters": nchk, "frames": fr, "against": against, "relerr_x": ex, "relerr_P": eP,
                       "ok": bool(ex <= 1e-9 and eP <= 1e-9)}}


def run_sweep(args, out):
    """--sweep: BASELINE config 4 (state-dimension sweep on one GPU); one JSON line per L."""
    import torch
    assert torch.cuda.is_available(), "bench.py --sweep needs a CUDA device (there is no CPU fallback)"
    from cv_monoslam_b200.slam import fp64_peak_tflops
    peak = fp64_peak_tflops(0)
    cfg = []
    for item in args.sweep.split(","):
        L = int(item.split(":")[0])
        n = 6 * L + 4
        npad = (n + 7) // 8 * 8
        per = 16.0 * npad * npad          # two np x np squares per filter dominate the footprint (<= 100 GB)
        B = int(item.split(":")[1]) if ":" in item else int(min(262144, max(1024, 2 ** int(np.log2(100e9 / per)))))
        cfg.append((L, B))
    for L, B in cfg:
        out.emit(json.dumps(sweep_point(L, B, args.steps, args.warmup, peak, args.sweep_parity_max_l, 0,
                                        gentle=L > args.sweep_gentle_above)))


class StdoutGuard:
    """Everything written to fd 1 while the benchmark runs (e.g. NCCL's version banner, written from C) goes to
    stderr; only the final JSON line reaches stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line: str):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        print(line, flush=True)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def main():
    with StdoutGuard() as out:
        run(out)


def run(out):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--filters", type=int, default=65536, help="filters per GPU (BASELINE config 3)")
    ap.add_argument("--landmarks", type=int, default=50)
    ap.add_argument("--unique", type=int, default=8, help="distinct synthetic worlds (priors) replicated over the batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--downdate-mode", type=int, default=0)
    ap.add_argument("--adversarial-frac", type=float, default=0.0,
                    help="force this share of the filters through the reference-order fallback every step (sets "
                         "SRUKF_FORCE_FALLBACK_PPM for the library): what the guard costs when it fires")
    ap.add_argument("--sweep", default=None, metavar="L[:B],...",
                    help="BASELINE config 4: state-dimension sweep on one GPU, e.g. 10,20,33,50,66,100 (B sized to ~60 GB "
                         "unless given); prints one JSON line per L with a parity spot check")
    ap.add_argument("--sweep-gentle-above", type=int, default=100,
                    help="sweep points with more landmarks than this use the gentler synthetic noise")
    ap.add_argument("--sweep-parity-max-l", type=int, default=100,
                    help="largest L whose sweep spot check runs the CPU oracle (beyond: the library's reference-arithmetic path)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.adversarial_frac > 0:
        os.environ["SRUKF_FORCE_FALLBACK_PPM"] = str(int(round(args.adversarial_frac * 1e6)))

    from cv_monoslam_b200 import dist
    if args.sweep:
        run_sweep(args, out)
        return
    if args.impl == "reference":
        ctx = dist.Ctx(int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), 0, None)
        run_reference(args, ctx, out)
        return

    import torch
    ctx = dist.init()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    dev = ctx.local_rank
    torch.cuda.set_device(dev)
    from cv_monoslam_b200 import CSLAMBatch, capi
    import synth
    B, L = args.filters, args.landmarks
    n = 6 * L + 4
    ntri = n * (n + 1) // 2
    W, K = args.warmup, args.steps
    T = 2 * (W + K)                       # device-resident pass, then the end-to-end pass continues the run
    first = ctx.rank * B                  # weak scaling: global filter ids of this rank
    sc = synth.make_scenario(L, B, T, unique=args.unique, first_filter=first, dense_state=False)

    g = CSLAMBatch(B, L, capi.default_params(downdate_mode=args.downdate_mode), device=dev)
    load_priors(g, sc, B)   # priors: upload the distinct worlds, replicate on the device
    stream = torch.cuda.ExternalStream(g.stream(), device=dev)

    # ---- pass 1: inputs resident in HBM ----------------------------------------------------------------
    half = W + K
    d_u = torch.from_numpy(sc.u[:half]).cuda()
    d_z = torch.from_numpy(sc.z[:half]).cuda()
    d_m = torch.from_numpy(sc.matched[:half]).cuda()
    torch.cuda.synchronize()
    for s in range(W):
        g.SLAM_dev(d_u[s].data_ptr(), d_z[s].data_ptr(), d_m[s].data_ptr())
    g.sync()
    g.set_profiling(True)
    launches0 = g.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(ctx)
    torch.cuda.synchronize()
    sampler = ClockSampler(dev)
    e0.record(stream)
    for s in range(W, W + K):
        g.SLAM_dev(d_u[s].data_ptr(), d_z[s].data_ptr(), d_m[s].data_ptr())
    g.sync()          # includes the side stream of the guard's fallback: the timed region ends when ALL the work is done
    e1.record(stream)
    g.sync()
    torch.cuda.synchronize()
    dist.barrier(ctx)
    clocks = sampler.stop()
    ms_total = dist.max_over_ranks(ctx, e0.elapsed_time(e1))
    launches = g.launch_count() - launches0
    kms, kcnt = g.kernel_times()
    g.set_profiling(False)
    del d_u, d_z, d_m

    # ---- pass 2: end to end through the host-pointer C ABI ---------------------------------------------
    hu = torch.from_numpy(sc.u[half:]).pin_memory()
    hz = torch.from_numpy(sc.z[half:]).pin_memory()
    hm = torch.from_numpy(sc.matched[half:]).pin_memory()
    hx = torch.empty((B, n), dtype=torch.float64).pin_memory()
    hx_np = hx.numpy()
    lib, hnd = g._lib, g._h
    def e2e_step(s):
        # inputs of this frame travel on the library's copy stream while the previous frame computes; the read-back of
        # m_X_k of this frame overlaps the next one (srukf_get_x_async) -- both inside the timed region
        capi.check(lib.srukf_step(hnd, hu[s].data_ptr(), hz[s].data_ptr(), hm[s].data_ptr()))
        capi.check(lib.srukf_get_x_async(hnd, hx.data_ptr()))
    for s in range(W):
        e2e_step(s)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(ctx)
    torch.cuda.synchronize()
    f0.record(stream)
    for s in range(W, W + K):
        e2e_step(s)
    g.sync()          # copy / read-back / fallback streams have drained
    f1.record(stream)
    g.sync()
    torch.cuda.synchronize()
    dist.barrier(ctx)
    ms_e2e = dist.max_over_ranks(ctx, f0.elapsed_time(f1))
    h2d = B * (3 * 8 + 2 * L * 8 + L)
    d2h = B * n * 8

    # ---- statistics: the system's only collective (NCCL all-reduce of 8 doubles) --------------------------
    part = g.stats(sc.truth[T - 1])
    tot = dist.allreduce_stats(ctx, part)
    stats = dist.summarise(tot)
    stats["finite_state"] = bool(np.isfinite(hx_np).all())
    flags = g.flags()
    stats["flag_or"] = int(np.bitwise_or.reduce(flags))
    stats["n_fallback"] = int(((flags & 32) != 0).sum())   # filters that went through the reference-order fallback at least once

    # ---- CPU baseline (rank 0, N == 1 only) ---------------------------------------------------------------
    cpu = None
    if ctx.rank == 0 and ctx.world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        kind, t1 = cpu_reference_run(L, 1, 2 if L >= 40 else 20, 0, 1)                 # one core, as the reference runs
        kind, tn = cpu_reference_run(L, cores, 2 if L >= 40 else 20, 0, cores)         # every core, one filter each
        cpu = {"value": cores * len(tn) / sum(tn), "unit": "filter-steps/s", "cores": cores, "kind": kind,
               "sample": f"{cores} filters (one per host thread) x {len(tn)} steps at L={L}: {CPU_KIND_TEXT[kind]}, "
                         f"{sum(tn):.1f} s",
               "one_core": {"value": len(t1) / sum(t1), "cores": 1, "seconds_per_filter_step": sum(t1) / len(t1),
                            "sample": f"1 filter x {len(t1)} steps"}}

    peak = FP64_PEAK_FALLBACK_TFLOPS
    peak_source = "fallback: profiles/r01_fp64_peak.json (tools/fp64_peak.cu, DMMA m8n8k4)"
    try:
        from cv_monoslam_b200.slam import fp64_peak_tflops
        peak = fp64_peak_tflops(dev)
        peak_source = ("measured in this run: srukf_fp64_peak (back-to-back DMMA m8n8k4, 8 warps x 4 CTAs per SM, best of 5) "
                       "right after the timed region; MEASURED_PEAKS.json has no FP64 entry")
    except Exception as e:   # noqa: BLE001
        peak_source += f" ({e})"
    hbm_peak = float(measured_peaks().get("hbm_gbs", 0.0)) or None

    if ctx.rank == 0:
        total_steps = float(B) * ctx.world * K
        value = total_steps / (ms_total * 1e-3)
        wd = flops_downdate(n, L)
        per_launch_filters = B * K / max(int(kcnt[2]), 1)
        t_dd = kms[2] / max(int(kcnt[2]), 1) * 1e-3
        achieved = wd * per_launch_filters / t_dd / 1e12 if t_dd > 0 else 0.0
        w_total = wd + flops_gain(n, L) + flops_predict(n, L)
        traffic = committed_traffic("k_update", L) if args.downdate_mode == 0 else None
        line = {
            "metric": "srukf_filter_steps_per_sec", "value": value, "unit": "filter-steps/s",
            "n_gpus": ctx.world, "steps": K, "warmup": W, "ms_per_step": ms_total / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(B, L, int(sc.meta["unique"]), ctx.world, args.downdate_mode),
            "clocks": clocks,
            "e2e": {"value": total_steps / (ms_e2e * 1e-3), "unit": "filter-steps/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / K},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "pipe": "fp64 (DFMA/DMMA share one pipe)", "kernel": "k_update",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak,
                         "traffic": (traffic[0] * per_launch_filters) if traffic else None,
                         "traffic_source": traffic[1] if traffic else None,
                         "peak_source": peak_source,
                         "flops_per_filter_step_kernel": wd, "flops_per_filter_step_all": w_total,
                         "whole_step_frac": value / ctx.world * w_total / (peak * 1e12),
                         "kernel_ms": {"k_predict": float(kms[0]), "k_gain": float(kms[1]), "k_update": float(kms[2])},
                         "kernel_launches": [int(c) for c in kcnt],
                         "algorithmic_bytes_per_filter_step": algorithmic_bytes(n, L),
                         "hbm_peak_gbs": hbm_peak,
                         "hbm_frac": (value / ctx.world * algorithmic_bytes(n, L) / (hbm_peak * 1e9)) if hbm_peak else None},
            "cpu_baseline": cpu,
            "stats": stats,
        }
        out.emit(json.dumps(line))
    g.close()
    dist.finalize(ctx)


if __name__ == "__main__":
    main()
