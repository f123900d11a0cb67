#!/usr/bin/env python
"""GPU-side (debug build with -DOGJK_PHASE_CLOCKS): average SM cycles per GJK iteration spent in the
two support scans vs the rest of the step (tests + sub-solver), and in the witness phase, per workload
and lanes-per-pair.  Usage: make -C opengjk_b200/csrc EXTRA=-DOGJK_PHASE_CLOCKS && python tools/phase_clocks.py"""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS
from opengjk_b200 import api, _lib
eng = api.Engine(0)
L = _lib.lib()
dev = torch.device("cuda:0")
CASES = (("hulls1k_f32", 32, 100000, 0), ("hulls1k_f32", 32, 100000, 256), ("hulls1k10k_f32", 32, 20000, 0),
         ("hulls1k10k_f32", 32, 20000, 256), ("hulls1k_f32", 132, 100000, 0), ("hulls64_f64", 2, 500000, 0),
         ("hulls64_f64", 104, 500000, 0), ("hulls64_f32", 2, 500000, 0))
if len(sys.argv) > 1:
    CASES = tuple(c for c in CASES if c[0] in sys.argv[1:])
for name, lanes, pairs, amin in CASES:
    desc, dt, maker, _ = WORKLOADS[name][:4]
    w = maker(pairs, 42)
    prec = "f64" if dt == np.float64 else "f32"
    sdt = api.SIMPLEX_F64 if prec == "f64" else api.SIMPLEX_F32
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    eng.set_accel_min(amin)
    store = eng.store_create(w.coords, w.off, w.cnt)
    pa, pb = t(w.pa), t(w.pb)
    simp = torch.empty(pairs * sdt.itemsize, dtype=torch.uint8, device=dev)
    dist = torch.empty(pairs, dtype=t(w.coords[:1]).dtype, device=dev)
    eng.set_classes([(2048, lanes)] if lanes < 100 else [(600 if "1k" in name else 40, lanes), (2048, lanes - 100)])
    for cap in ("1", "16"):
        os.environ["OGJK_GRID_CAP"] = cap
        out = (C.c_ulonglong * 8)()
        L.ogjk_debug_phase_clocks(out)
        eng.store_gjk(store, pa, store, pb, pairs, simp, dist, 0)
        L.ogjk_debug_phase_clocks(out)
        steps, npairs = max(out[3], 1), max(out[4], 1)
        print(f"{name:12s} lanes={lanes:2d} tables>={amin:<3d} blocks/SM cap={cap:>2s}: support {out[0]/steps:7.0f} cyc/iter  rest-of-step {out[1]/steps:7.0f} cyc/iter  "
              f"witness {out[2]/npairs:6.0f} cyc/pair  iters/pair {steps/npairs:.2f}  table trips/iter {out[6]/steps:.2f} live-at-trip {out[5]/max(out[6],1):.1f}  whole grab {out[7]/max(out[4]*lanes/32 if lanes<=32 else out[4],1):.0f} cyc", flush=True)
    store.close()
