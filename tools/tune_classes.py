#!/usr/bin/env python
"""GPU-side sweep: time ogjk_store_gjk_* with every uniform lanes-per-pair choice
on the bench workloads (device-resident store, CUDA events).  Prints one line per
(workload, lanes).  Usage: python tools/tune_classes.py [workload ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS  # noqa: E402
from opengjk_b200 import api, workloads as W  # noqa: E402

names = sys.argv[1:] or ["hulls64_f64", "hulls64_f32", "boxes_f32", "hulls1k_f32"]
WORKLOADS["hulls256_f32"] = ("64-256 vertices fp32", np.float32, lambda n, seed: W.random_hulls(n, 64, 256, seed, 4.0, np.float32), 300_000)
WORKLOADS["hulls256_f64"] = ("64-256 vertices fp64", np.float64, lambda n, seed: W.random_hulls(n, 64, 256, seed, 4.0, np.float64), 300_000)
WORKLOADS["hulls600_f32"] = ("256-600 vertices fp32", np.float32, lambda n, seed: W.random_hulls(n, 256, 600, seed, 4.0, np.float32), 100_000)
eng = api.Engine(0)
dev = torch.device("cuda:0")
for name in names:
    desc, dt, maker, pairs = WORKLOADS[name]
    pairs = min(pairs, 2_000_000)
    w = maker(pairs, 42)
    prec = "f64" if dt == np.float64 else "f32"
    sdt = api.SIMPLEX_F64 if prec == "f64" else api.SIMPLEX_F32
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    coords, off, cnt, pa, pb = t(w.coords.reshape(-1)), t(w.off), t(w.cnt), t(w.pa), t(w.pb)
    simp = torch.empty(w.npairs * sdt.itemsize, dtype=torch.uint8, device=dev)
    dist = torch.empty(w.npairs, dtype=coords.dtype, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    store = eng.store_create(coords, off, cnt, st)
    tables = {"small+g": None}
    small_ok = int(w.cnt.max()) <= 8
    kmax = int((((w.cnt[w.pa] + 3) // 4 * 4 + (w.cnt[w.pb] + 3) // 4 * 4) // 4).max())
    variants = ([0, 200] if small_ok else []) + [1, 2, 4, 8, 16, 32]
    if os.environ.get("TUNE_ASSIST") and kmax > 40:
        variants += [302, 304, 308, 316]
    if 40 < kmax < 2048:
        variants += [502, 504, 508, 516]
    if os.environ.get("TUNE_PERSIST"):
        variants += [201, 202, 204, 208, 232]
    if os.environ.get("TUNE_STAGED"):
        variants += [102, 104, 108, 132]
        if 12 < kmax <= 40:
            variants += [1102, 1104, 1108]  # three slot sizes
    for g in variants:
        if g == 0:
            table = [(4, 0), (2048, 1)]
        elif g > 1000:
            gg = g - 1000
            table = [(12, gg), (20, gg), (kmax, gg), (2048, gg - 100)]
        elif g == 200:
            table = [(4, 200), (2048, 201)]
        elif g > 500:
            table = [(kmax, g), (2048, 32)]  # CTA-per-pair: shared memory sized for the largest pair
        elif g > 200:
            table = [(2048, g)]  # persistent-lane (2xx) and scan-assist (3xx) flavours
        elif g >= 100:
            table = [(kmax, g), (2048, g - 100)]
        else:
            table = [(2048, g)]
        if kmax >= 2048 and g >= 100:
            continue
        eng.set_classes(table)
        for _ in range(3):
            eng.store_gjk(store, pa, store, pb, w.npairs, simp, dist, st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            eng.store_gjk(store, pa, store, pb, w.npairs, simp, dist, st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"{name:14s} lanes={g:4d}  {ms:8.3f} ms  {w.npairs / ms / 1e3:10.1f} Mpairs/s", flush=True)
    store.close()
