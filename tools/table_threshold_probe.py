#!/usr/bin/env python
"""GPU-side: resident-store throughput with and without cluster tables for hull sizes around the
table threshold (default classes).  Usage: python tools/table_threshold_probe.py"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opengjk_b200 import api, workloads as W
eng = api.Engine(0); dev = torch.device("cuda:0")
for lo, hi, pairs, dt in ((256, 400, 200000, np.float32), (256, 600, 200000, np.float32), (400, 800, 200000, np.float32),
                          (600, 1000, 200000, np.float32), (256, 600, 100000, np.float64), (400, 800, 100000, np.float64), (1000, 1000, 100000, np.float64)):
    w = W.random_hulls(pairs, lo, hi, 42, 4.0, dt)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    sdt = api.SIMPLEX_F64 if dt == np.float64 else api.SIMPLEX_F32
    res = []
    for amin in (0, 256, 512):
        eng.set_accel_min(amin)
        store = eng.store_create(w.coords, w.off, w.cnt)
        pa, pb = t(w.pa), t(w.pb)
        simp = torch.empty(w.npairs * sdt.itemsize, dtype=torch.uint8, device=dev)
        dist = torch.empty(w.npairs, dtype=torch.float64 if dt == np.float64 else torch.float32, device=dev)
        for _ in range(3): eng.store_gjk(store, pa, store, pb, w.npairs, simp, dist, 0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(5): eng.store_gjk(store, pa, store, pb, w.npairs, simp, dist, 0)
        b.record(); torch.cuda.synchronize()
        res.append(w.npairs / (a.elapsed_time(b) / 5) / 1e3)
        store.close()
    print(f"hulls {lo}-{hi} {np.dtype(dt).name}: plain {res[0]:.1f}  tables>=256 {res[1]:.1f} ({res[1]/res[0]:.2f}x)  tables>=512 {res[2]:.1f} ({res[2]/res[0]:.2f}x) Mpairs/s", flush=True)
