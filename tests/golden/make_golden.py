"""Generates tests/golden/*.npz: seeded inputs and the C oracle's outputs (literal mode).

The reference ships no golden vectors (PARITY UNPINNED, see oracle/srukf_oracle.h); these fixtures pin
the oracle against regressions and give the GPU tests committed, reference-free expectations.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import oracle as O  # noqa: E402
import synth  # noqa: E402

CASES = {
    # name: (L, B, steps, match_prob)
    "L3_B4_s6": (3, 4, 6, 1.0),
    "L8_B4_s8": (8, 4, 8, 1.0),
    "L8_B4_s6_ragged": (8, 4, 6, 0.6),
    "L20_B2_s4": (20, 2, 4, 1.0),
}


def make(name, L, B, steps, match_prob):
    sc = synth.make_scenario(L, B, steps, match_prob=match_prob)
    p = O.default_params(downdate_mode=0)  # literal: S^T S re-formed per U column
    x, S = sc.x0.copy(), sc.S0.copy()
    n = 6 * L + 4
    xs = np.empty((steps, B, n))
    Ps = np.empty((steps, B, n, n))
    for s in range(steps):
        O.batch_step(p, x, S, sc.u[s:s + 1], sc.z[s:s + 1], sc.matched[s:s + 1], 8)
        xs[s] = x
        Ps[s] = np.einsum("bki,bkj->bij", S, S)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), L=L, B=B, steps=steps, x0=sc.x0, S0=sc.S0, u=sc.u,
                        z=sc.z, matched=sc.matched, truth=sc.truth, x=xs, P=Ps)
    print(name, "written; |x| max", np.abs(xs).max(), "tr P last", np.trace(Ps[-1, 0]))


if __name__ == "__main__":
    for k, v in CASES.items():
        make(k, *v)
