#!/usr/bin/env python
"""Same-box comparison with the reference's own CUDA kernels (gpu/opengjk_gpu.cu built
unmodified into oracle/_ref by `make -C oracle ref_gpu`): kernel-only time through its
mid-level API (device arrays built once) vs ogjk_store_gjk_* / ogjk_store_gjk_epa_* on the
same pairs.  Prints one JSON line per workload.  Test/bench infrastructure only."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) >= 3 and sys.argv[1] == "--npz":
    # bench.py's baseline leg: time ONLY the reference's kernels on the arrays of an .npz file, in this process
    # (a fault inside the reference library must not take the bench down).  argv: --npz file [epa]
    from oracle.pyoracle import ReferenceGPU
    z = np.load(sys.argv[2])
    epa = len(sys.argv) > 3 and sys.argv[3] == "epa"
    prec = "f64" if z["coords"].dtype == np.float64 else "f32"
    if not ReferenceGPU.available(prec):
        print(json.dumps({"unavailable": "oracle/_ref reference CUDA library not shipped"}))
        sys.exit(0)
    fd = os.dup(1)
    os.dup2(2, 1)  # the reference prints diagnostics on stdout
    t = ReferenceGPU(prec).time_kernels(z["coords"], z["off"], z["cnt"], z["pa"], z["pb"], reps=3, epa=epa)
    os.dup2(fd, 1)
    print(json.dumps({"ms": t["gjk_epa_ms"] if epa else t["gjk_ms"], "pairs": int(z["pa"].shape[0])}), flush=True)
    sys.exit(0)

import torch

from opengjk_b200 import api, workloads as W  # noqa: E402
from oracle.pyoracle import ReferenceGPU  # noqa: E402

CASES = [
    ("hulls64_f32", lambda: W.mixed_hulls(200_000, 8, 64, 42, np.float32), False),
    ("hulls64_f64", lambda: W.mixed_hulls(200_000, 8, 64, 42, np.float64), False),
    ("boxes_f32", lambda: W.boxes(400_000, True, 42, dtype=np.float32), False),
    ("hulls1k_f32", lambda: W.large_hull_pool(20_000, 4096, 1000, 1000, 42, dtype=np.float32), False),
    ("epa_f32", lambda: W.random_hulls(100_000, 8, 64, 42, 1.0, np.float32, 0.5), True),
]
eng = api.Engine(0)
dev = torch.device("cuda:0")
for name, make, epa in CASES:
    w = make()
    prec = "f64" if w.coords.dtype == np.float64 else "f32"
    if not ReferenceGPU.available(prec):
        print(json.dumps({"workload": name, "unavailable": "oracle/_ref reference CUDA library not shipped"}))
        continue
    ref = ReferenceGPU(prec).time_kernels(w.coords, w.off, w.cnt, w.pa, w.pb, reps=5, epa=epa)
    sdt = api.SIMPLEX_F64 if prec == "f64" else api.SIMPLEX_F32
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    store = eng.store_create(w.coords, w.off, w.cnt)
    pa, pb = t(w.pa), t(w.pb)
    n = w.npairs
    simp = torch.empty(n * sdt.itemsize, dtype=torch.uint8, device=dev)
    dist = torch.empty(n, dtype=t(w.coords[:1]).dtype, device=dev)
    w1, w2, nr = (torch.empty(n * 3, dtype=dist.dtype, device=dev) for _ in range(3))
    st = torch.cuda.current_stream().cuda_stream
    run = (lambda: eng.store_gjk_epa(store, pa, store, pb, n, simp, dist, w1, w2, nr, st)) if epa else (
        lambda: eng.store_gjk(store, pa, store, pb, n, simp, dist, st))
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        run()
    e1.record()
    torch.cuda.synchronize()
    ours = e0.elapsed_time(e1) / 5
    theirs = ref["gjk_epa_ms"] if epa else ref["gjk_ms"]
    print(json.dumps({"workload": name, "pairs": n, "reference_cuda_ms": theirs, "ours_ms": ours,
                      "reference_cuda_pairs_per_s": n / theirs * 1e3, "ours_pairs_per_s": n / ours * 1e3,
                      "speedup": theirs / ours, "note": "kernel-only, device-resident inputs on both sides"}), flush=True)
    store.close()
