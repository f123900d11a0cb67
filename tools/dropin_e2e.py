#!/usr/bin/env python
"""GPU-side: the SAME reference call, compute_minimum_distance(n, bd1, bd2, simplices, distances) with arrays of
gkPolytope structs that point at host vertex arrays, timed end to end (wall clock, host buffers in and out) through
(a) this repo's reference-named library opengjk_b200/libopenGJK_GPU.so and (b) the reference's own CUDA library built
unmodified into oracle/_ref (when shipped).  What a user who swaps the .so sees.  One JSON line per run."""
import ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opengjk_b200 import workloads as W
from oracle.pyoracle import polytope_dtype, simplex_dtype

def bodies(w, which, prec):
    arr = np.zeros(which.shape[0], dtype=polytope_dtype(prec))
    arr["numpoints"] = w.cnt[which]
    arr["coord"] = w.coords.ctypes.data + w.off[which].astype(np.uint64) * np.uint64(3 * w.coords.itemsize)
    return arr

p = lambda a: a.ctypes.data_as(C.c_void_p)
LIBS = [("ours", {"f32": "opengjk_b200/libopenGJK_GPU.so", "f64": "opengjk_b200/libopenGJK_GPU_f64.so"}),
        ("reference", {"f32": "oracle/_ref/libopenGJK_GPU_ref_f32.so", "f64": "oracle/_ref/libopenGJK_GPU_ref_f64.so"})]
for name, maker, n in (("hulls 8-64 v", lambda n, dt: W.mixed_hulls(n, dtype=dt), 200_000),
                       ("hulls 8-64 v (config 2 size; reference library skipped: 13 s per call)", lambda n, dt: W.mixed_hulls(n, dtype=dt), 1_000_000),
                       ("boxes", lambda n, dt: W.boxes(n, True, 42, dtype=dt), 1_000_000)):
    for prec, dt in (("f32", np.float32), ("f64", np.float64)):
        w = maker(n, dt)
        b1, b2 = bodies(w, w.pa, prec), bodies(w, w.pb, prec)
        res = {}
        for who, paths in LIBS:
            path = os.path.join(ROOT, paths[prec])
            if not os.path.exists(path) or (who == "reference" and "skipped" in name):
                continue
            lib = C.CDLL(path)
            simp = np.zeros(n, dtype=simplex_dtype(prec)); dist = np.zeros(n, dtype=dt)
            lib.compute_minimum_distance(C.c_int(n), p(b1), p(b2), p(simp), p(dist))   # warm-up (context, allocations)
            reps = 3 if who == "ours" else 1
            t0 = time.perf_counter()
            for _ in range(reps):
                lib.compute_minimum_distance(C.c_int(n), p(b1), p(b2), p(simp), p(dist))
            res[who] = n * reps / (time.perf_counter() - t0)
            res[who + "_checksum"] = float(np.nansum(dist))
        if "skipped" in name:
            # the flat-pool entry point with plain (pageable) numpy arrays, as a ctypes binding passes them
            from opengjk_b200 import api
            eng = api.Engine(0)
            pools = w.two_pools()
            got = eng.gjk_host_pools(*pools)
            t0 = time.perf_counter()
            for _ in range(3):
                got = eng.gjk_host_pools(*pools)
            res["ours_flat_pageable"] = n * 3 / (time.perf_counter() - t0)
            res["ours_flat_pageable_checksum"] = float(np.nansum(got[0] if isinstance(got, tuple) else got["distances"]))
            del eng
        out = {"workload": name, "prec": prec, "pairs": n, "pairs_per_s": {k: v for k, v in res.items() if not k.endswith("checksum")}}
        if "ours_flat_pageable" in res:
            out["distance_sums"] = [res["ours_checksum"], res["ours_flat_pageable_checksum"]]
        if "ours" in res and "reference" in res:
            out["speedup"] = res["ours"] / res["reference"]
            out["distance_sums"] = [res["ours_checksum"], res["reference_checksum"]]
        print(json.dumps(out), flush=True)
