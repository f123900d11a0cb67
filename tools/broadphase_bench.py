#!/usr/bin/env python
"""GPU-side timing of the broad-phase kernels (CUDA events, after warm-up): box pass in GB/s of
store bytes read against the measured HBM peak, candidate culling in pairs/s, all-pairs in box
tests/s.  Prints one JSON line per kernel.  Usage: python tools/broadphase_bench.py"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import peaks
from opengjk_b200 import api, workloads as W

eng = api.Engine(0)
dev = torch.device("cuda:0")
hbm, src = peaks()
st = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for name, w in (("hulls 8-64 v fp64, 2 M bodies", W.mixed_hulls(1_000_000, dtype=np.float64)),
                ("hulls 1000 v fp32, 32 k bodies", W.large_hull_pool(1000, 32768, 1000, 1000, 1, dtype=np.float32))):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    eng.set_accel_min(0)
    store = eng.store_create(t(w.coords.reshape(-1)), t(w.off), t(w.cnt), st)
    P = store.num_polytopes
    boxes = torch.empty((P, 6), dtype=torch.float64 if w.coords.dtype == np.float64 else torch.float32, device=dev)
    ms = timed(lambda: eng.store_aabbs(store, boxes, st))
    nbytes = int(((w.cnt + 3) // 4 * 4).sum()) * 3 * w.coords.itemsize + P * (12 + 6 * w.coords.itemsize)
    print(json.dumps({"kernel": "aabb_kernel", "input": name, "ms": ms, "GB/s": nbytes / ms / 1e6,
                      "frac_of_hbm_peak": nbytes / ms / 1e6 / hbm, "peak_source": src}), flush=True)
    prec = "f64" if w.coords.dtype == np.float64 else "f32"
    n = 20_000_000
    g = torch.Generator(device=dev); g.manual_seed(1)
    i1 = torch.randint(0, P, (n,), device=dev, dtype=torch.int32, generator=g)
    i2 = torch.randint(0, P, (n,), device=dev, dtype=torch.int32, generator=g)
    o1, o2 = torch.empty_like(i1), torch.empty_like(i2)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    ms = timed(lambda: eng.pairs_cull(prec, boxes, boxes, n, i1, i2, 0.0, n, o1, o2, cnt, st), 10)
    print(json.dumps({"kernel": "cull_count+scan+scatter", "input": name, "candidates": n, "kept": int(cnt.item()),
                      "ms": ms, "Mpairs/s": n / ms / 1e3}), flush=True)
    m = min(P, 60_000)
    ms = timed(lambda: eng.pairs_all(prec, boxes, m, 0.0, n, o1, o2, cnt, st), 3)
    print(json.dumps({"kernel": "allpairs (count+scan+write)", "input": f"first {m} bodies of " + name, "pairs_found": int(cnt.item()),
                      "ms": ms, "G box tests/s": m * (m - 1) / 2 * 2 / ms / 1e6}), flush=True)
    store.close()
