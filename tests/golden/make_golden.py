"""Generates tests/golden/*.npz: seeded inputs and the outputs of the REFERENCE'S OWN CODE on them.

The reference ships no golden vectors.  These are produced by oracle/_ref/libsrukf_ref.so -- the bodies of the CSLAM
member functions on the path, extracted verbatim from /root/reference/MonoSLAM/SLAM.cpp at build time and compiled
against oracle/ref_shim/ (see oracle/ref.py) -- so the fixtures are reference outputs, not outputs of this repo's
restatement.  tests/test_oracle.py checks the C oracle against them, tests/test_gpu_parity.py the CUDA path, and
tests/test_ref_pin.py re-runs the reference on them where /root/reference (or the prebuilt library) is available.

Stored per case: x0 [B,n], S0 [B,n,n]; per step the two odometry poses handed to predictMotion (`odo`
[steps,B,2,3]), the control the reference derived from them (`u`, SLAM.cpp:1446-1450), z, matched, and the
reference's m_X_k (`x`) and m_S_k^T m_S_k (`P`) after KalmanUpdate.

Run from the repo root (needs /root/reference):  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref as R  # noqa: E402
import synth  # noqa: E402

CASES = {
    # name: (L, B, steps, match_prob, weight_type)
    "L3_B4_s6": (3, 4, 6, 1.0, 0),
    "L8_B4_s8": (8, 4, 8, 1.0, 0),
    "L8_B4_s6_ragged": (8, 4, 6, 0.6, 0),
    "L20_B2_s4": (20, 2, 4, 1.0, 0),
    "L3_B2_s1_weights2": (3, 2, 1, 1.0, 1),   # wm0 ~ -1e6: ill-conditioned in the reference too, one frame
    "L5_B2_s3_weights3": (5, 2, 3, 1.0, 2),
    "L50_B1_s1": (50, 1, 1, 1.0, 0),
}


def make(name, L, B, steps, match_prob, weight_type):
    sc = synth.make_scenario(L, B, steps, match_prob=match_prob)
    n = 6 * L + 4
    xs = np.empty((steps, B, n))
    Ps = np.empty((steps, B, n, n))
    us = np.empty((steps, B, 3))
    odo = np.empty((steps, B, 2, 3))
    for b in range(B):
        r = R.Slam(weight_type)
        r.set_state(sc.x0[b], sc.S0[b])
        for s in range(steps):
            odo[s, b, 0], odo[s, b, 1] = R.control_to_odometry(sc.u[s, b])
            us[s, b], _ = r.predict_motion_odometry(odo[s, b, 0], odo[s, b, 1])
            r.predict_measurement()
            r.kalman_update(sc.z[s, b], sc.matched[s, b])
            x, S = r.get_state()
            xs[s, b] = x
            Ps[s, b] = S.T @ S
    assert np.abs(us - sc.u).max() < 1e-15
    np.savez_compressed(os.path.join(HERE, name + ".npz"), L=L, B=B, steps=steps, weight_type=weight_type,
                        source="reference: oracle/_ref/libsrukf_ref.so (SLAM.cpp bodies extracted verbatim)",
                        x0=sc.x0, S0=sc.S0, odo=odo, u=us, z=sc.z, matched=sc.matched, truth=sc.truth, x=xs, P=Ps)
    print(name, "written; |x| max", np.abs(xs).max(), "tr P last", np.trace(Ps[-1, 0]))


if __name__ == "__main__":
    for k, v in CASES.items():
        make(k, *v)
