"""Extracts the one pair of the default bench workload (mixed_hulls, seed 42, 1 M pairs) that
ends on the iteration cap (scalar/openGJK.c:1036-1042, k == mk) and stores its two hulls as
tests/golden/cap_pair.npz.  The configuration is robustly cap-prone: rigid motions and small
perturbations of it hit the cap in 60-100 % of the cases (workloads.cap_stress), which is how
the parity tests get thousands of pairs on that exit.  Run from the repo root:
    python tests/golden/make_cap_pair.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from opengjk_b200 import workloads as W  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402

if __name__ == "__main__":
    w = W.mixed_hulls(1_000_000, 8, 64, 42, np.float64)
    d, s, it = Oracle("f64").batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=os.cpu_count() or 1, want_iters=True)
    idx = np.nonzero(it >= 25)[0]
    assert idx.size >= 1, "no cap-hitting pair found"
    p = int(idx[0])
    a, b = w.polytope(w.pa[p]).copy(), w.polytope(w.pb[p]).copy()
    np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "cap_pair.npz"), a=a, b=b, pair=p, seed=42)
    print(f"pair {p}: {a.shape[0]} + {b.shape[0]} vertices, distance {d[p]:.17g}, {it[p]} iterations")
