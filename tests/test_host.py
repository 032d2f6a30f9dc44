"""CPU tests of the host-side logic: synthetic inputs, sharding, statistics reduction (gloo, world 2)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from cv_monoslam_b200 import dist
import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_scenario_shapes_and_determinism():
    a = synth.make_scenario(4, 6, 3, unique=2)
    b = synth.make_scenario(4, 6, 3, unique=2)
    n = 28
    assert a.x0.shape == (6, n) and a.S0.shape == (6, n, n)
    assert a.u.shape == (3, 6, 3) and a.z.shape == (3, 6, 4, 2) and a.matched.shape == (3, 6, 4)
    assert np.array_equal(a.z, b.z) and np.array_equal(a.S0, b.S0)
    assert np.array_equal(a.x0[0], a.x0[2]) and not np.array_equal(a.z[:, 0], a.z[:, 2])   # replicated world, own noise
    assert np.allclose(np.tril(a.S0[0], -1), 0) and np.all(np.diag(a.S0[0]) > 0)


def test_measurements_stay_inside_the_view_for_the_parity_horizon():
    sc = synth.make_scenario(20, 4, 100, unique=4)
    assert sc.z.min() > 20 and sc.z[..., 0].max() < 620 and sc.z[..., 1].max() < 460


def test_per_filter_inputs_do_not_depend_on_sharding():
    full = synth.make_scenario(3, 8, 2, unique=8)
    for rank in range(2):
        lo, hi = dist.shard_range(8, rank, 2)
        part = synth.make_scenario(3, hi - lo, 2, unique=8, first_filter=lo)
        assert np.array_equal(part.z, full.z[:, lo:hi]) and np.array_equal(part.u, full.u[:, lo:hi])
        assert np.array_equal(part.x0, full.x0[lo:hi])


def test_shard_range_covers_everything_once():
    for B, W in ((10, 3), (65536, 8), (7, 8)):
        seen = []
        for r in range(W):
            lo, hi = dist.shard_range(B, r, W)
            seen += list(range(lo, hi))
        assert seen == list(range(B))


def test_summarise_stats():
    s = dist.summarise(np.array([4.0, 9.0, 1.0, 30.0, 10.0, 0.0, 1.0, 0.0]))
    assert s["rmse_xy"] == pytest.approx(np.sqrt(1.3)) and s["nees"] == pytest.approx(3.0)
    assert s["filters"] == 10 and s["flag_gmw_modified"] == 1


def test_stats_allreduce_world2_gloo(tmp_path):
    """Two CPU ranks over gloo: the only collective of the system."""
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys, numpy as np\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from cv_monoslam_b200 import dist\n"
        "ctx = dist.init(backend='gloo')\n"
        "part = np.arange(8, dtype=np.float64) * (ctx.rank + 1)\n"
        "tot = dist.allreduce_stats(ctx, part)\n"
        "t = dist.max_over_ranks(ctx, 1.0 + ctx.rank)\n"
        "assert np.array_equal(tot, np.arange(8) * 3.0), tot\n"
        "assert t == 2.0\n"
        "lo, hi = dist.shard_range(10, ctx.rank, ctx.world)\n"
        "assert (lo, hi) == ((0, 5) if ctx.rank == 0 else (5, 10))\n"
        "dist.barrier(ctx)\n"
        "if ctx.rank == 0: print('OK')\n"
        "dist.finalize(ctx)\n")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "OK" in out.stdout


def test_bench_arms_share_one_config_and_traffic_comes_from_a_committed_profile(tmp_path, monkeypatch):
    """bench.py: both arms describe the workload with the same `config` object (the reference arm times a sample of it),
    and roofline.traffic is read from a committed ncu summary, never a constant"""
    import bench
    a = bench.config_dict(65536, 50, 8, 1, 0)
    b = bench.config_dict(65536, 50, 8, 1, 0)
    assert a == b and "65536 filters x 50 landmarks" in a["workload"] and a["state_dim"] == 304
    assert bench.algorithmic_bytes(304, 50) == 8.0 * (304 * 305 + 2 * 304 + 3 + 100)
    assert abs(bench.flops_downdate(304, 50) + bench.flops_gain(304, 50) + bench.flops_predict(304, 50) - 32.39256e6) < 1e3
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    assert bench.committed_traffic("k_update", 50) is None
    (tmp_path / "profiles").mkdir()
    (tmp_path / "profiles" / "r99_traffic.json").write_text(
        '{"landmarks": 50, "kernels": {"k_update": {"dram_bytes_per_filter": 123.0}}}')
    assert bench.committed_traffic("k_update", 50) == (123.0, "profiles/r99_traffic.json")
    assert bench.committed_traffic("k_update", 20) is None
