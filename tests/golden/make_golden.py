#!/usr/bin/env python
"""Regenerates the golden fixtures in tests/golden/ (run in the build container,
where /root/reference is mounted; the GPU box only ever reads the outputs).

1. subsolver_kat.json -- the reference's own known-answer vectors for hff1/2/3
   and S1D/S2D/S3D, extracted as DATA from its cmocka sources
   (scalar/tests/HFFtest_test.c, subalg_{1,2,3}simplex_test.c).  The C files set
   a global simplex field by field, call subalgorithm(), then assert v[] and
   nvrtx; the parser below replays those assignments.
2. gjk_golden_{f64,f32}.npz -- outputs (distance + full gkSimplex bytes) of the
   UNMODIFIED reference (oracle/_ref/libopengjk_ref_*.so) on seeded workloads
   from opengjk_b200.workloads; inputs are regenerated from the seeds by the
   tests and guarded by a checksum stored next to the outputs.

Usage: python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("OGJK_REFERENCE", "/root/reference")

NUM = r"([-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?)f?"


def parse_subalg(path):
    cases = []
    nvrtx = None
    vrtx = [[None] * 3 for _ in range(4)]
    cur = None
    func = None
    for ln, line in enumerate(open(path), 1):
        line = line.strip()
        m = re.match(r"(subalg_\w+)\(void\*\* state\)", line)
        if m:
            func = m.group(1)
        m = re.match(r"s\.nvrtx = (\d+);", line)
        if m:
            nvrtx = int(m.group(1))
            cur = None
            continue
        m = re.match(r"s\.vrtx\[(\d)\]\[(\d)\] = " + NUM + ";", line)
        if m:
            vrtx[int(m.group(1))][int(m.group(2))] = float(m.group(3))
            cur = None
            continue
        if line.startswith("subalgorithm(&s, v);"):
            assert nvrtx is not None, f"{path}:{ln} nvrtx unknown"
            pts = vrtx[:nvrtx]
            assert all(c is not None for p in pts for c in p), f"{path}:{ln} stale simplex"
            cur = {"src": f"{os.path.basename(path)}:{ln}", "func": func, "vrtx": [list(p) for p in pts],
                   "v": [None, None, None], "nvrtx": None}
            cases.append(cur)
            vrtx = [[None] * 3 for _ in range(4)]  # the call rewrites the global simplex
            nvrtx = None
            continue
        m = re.match(r"assert_float_equal\(v\[(\d)\], " + NUM + r", FLOAT_TOL\);", line)
        if m and cur is not None:
            cur["v"][int(m.group(1))] = float(m.group(2))
            continue
        m = re.match(r"assert_int_equal\(s\.nvrtx, (\d)\);", line)
        if m and cur is not None:
            cur["nvrtx"] = int(m.group(1))
            nvrtx = cur["nvrtx"]  # the global simplex keeps the solver's result: some cases rely on it
    return cases


def parse_hff(path):
    cases = []
    vec = {k: [None] * 3 for k in "pqrs"}
    for ln, line in enumerate(open(path), 1):
        line = line.strip()
        m = re.match(r"([pqrs])\[(\d)\] = " + NUM + ";", line)
        if m:
            vec[m.group(1)][int(m.group(2))] = float(m.group(3))
            continue
        m = re.match(r"assert_int_equal\(hff(\d)\(([pqrs, ]+)\), (\d)\);", line)
        if m:
            args = [a.strip() for a in m.group(2).split(",")]
            cases.append({"src": f"{os.path.basename(path)}:{ln}", "which": int(m.group(1)),
                          "args": [list(vec[a]) for a in args], "expect": int(m.group(3))})
    return cases


def make_kat():
    t = os.path.join(REF, "scalar", "tests")
    out = {"hff": parse_hff(os.path.join(t, "HFFtest_test.c")), "sub": []}
    for k in (1, 2, 3):
        out["sub"] += parse_subalg(os.path.join(t, f"subalg_{k}simplex_test.c"))
    with open(os.path.join(HERE, "subsolver_kat.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("subsolver_kat.json:", len(out["hff"]), "hff cases,", len(out["sub"]), "sub-solver cases")


def golden_workloads(dtype):
    """The seeded workloads the golden outputs were produced on (tests import this)."""
    from opengjk_b200 import workloads as W
    return [
        W.from_bodies([W.USER_P], [W.USER_Q], dtype, "userPQ"),
        W.mixed_hulls(3000, seed=42, dtype=dtype),
        W.boxes(3000, True, 42, dtype=dtype),
        W.boxes(3000, False, 42, dtype=dtype),
        W.tiny_bodies(40, 12345, 1.0, dtype),
        W.tiny_bodies(40, 12346, 1e4, dtype),
        W.random_hulls(2000, 8, 64, 7, 1.0, dtype, 0.5),
        W.large_hull_pool(300, 32, 200, 1500, 42, dtype=dtype),
    ]


def checksum(w):
    h = hashlib.sha256()
    for a in (w.coords, w.off, w.cnt, w.pa, w.pb):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def make_outputs():
    from oracle.pyoracle import Reference
    for prec, dt in (("f64", np.float64), ("f32", np.float32)):
        ref = Reference(prec)
        blob = {}
        for w in golden_workloads(dt):
            d, s = ref.batch(w.coords, w.off, w.cnt, w.pa, w.pb)
            blob[w.name + "/dist"] = d
            blob[w.name + "/simplex"] = s.view(np.uint8).reshape(len(s), -1)
            blob[w.name + "/sha"] = np.frombuffer(checksum(w).encode(), dtype=np.uint8)
            print(prec, w.name, w.npairs, "pairs")
        np.savez_compressed(os.path.join(HERE, f"gjk_golden_{prec}.npz"), **blob)


if __name__ == "__main__":
    make_kat()
    make_outputs()
