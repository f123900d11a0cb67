#!/usr/bin/env python
"""Generates tests/golden/epa_ref_cuda_{f32,f64}.npz ON A GPU BOX: outputs of the reference's
own CUDA GJK+EPA (gpu/opengjk_gpu.cu compiled unmodified into oracle/_ref by `make -C oracle
ref_gpu`, once with its default flags and once with -fmad=false) on seeded workloads.  The
reference has no CPU EPA and no numeric EPA test, so these files are the pin for
oracle/epa_oracle.c: tests/test_oracle_epa_golden.py compares the CPU restatement with them.

Usage (GPU box): python tests/golden/make_epa_golden.py gpurun_out/   (then copy into tests/golden/)"""
import hashlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from opengjk_b200 import workloads as W
from oracle.pyoracle import ReferenceGPU


def epa_workloads(dt):
    """(name, Workload): penetrating hull pairs (config 5), overlapping boxes, a separated slice."""
    return [("hulls8-64_overlap", W.random_hulls(3000, 8, 64, 23, 1.0, dt, 0.5)),
            ("boxes_overlap", W.boxes(1500, True, 24, 1.5, dt)),
            ("cubes_aligned_overlap", W.boxes(1000, False, 26, 1.5, dt)),
            ("hulls_mixed", W.random_hulls(1500, 8, 64, 25, 4.0, dt, 0.5))]


def checksum(w):
    h = hashlib.sha256()
    for a in (w.coords, w.off, w.cnt, w.pa, w.pb):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def _one(prec, variant, name):
    """Child process: one library, one workload (a CUDA fault in the reference kernel is sticky
    for the process, so every run is isolated)."""
    dt = np.float32 if prec == "f32" else np.float64
    w = dict(epa_workloads(dt))[name]
    ref = ReferenceGPU(prec, variant)
    d, s, w1, w2 = ref.gjk(w.coords, w.off, w.cnt, w.pa, w.pb, epa=True)
    d2, _, _, _ = ref.gjk(w.coords, w.off, w.cnt, w.pa, w.pb, epa=True)   # run-to-run reproducibility
    # the two stages separately: the reference's own GPU GJK results, then its EPA on exactly those
    gd, gs = ref.gjk(w.coords, w.off, w.cnt, w.pa, w.pb)
    ed, ew1, ew2, enr = ref.epa(w.coords, w.off, w.cnt, w.pa, w.pb, gs, gd)
    import ctypes
    err = ctypes.CDLL("libcudart.so").cudaDeviceSynchronize()
    np.savez(sys.argv[5], dist=d, w1=w1, w2=w2, repeatable=np.array(d.tobytes() == d2.tobytes()), cuda_error=np.array(err),
             gjk_dist=gd, gjk_simplex=gs.view(np.uint8), epa_dist=ed, epa_w1=ew1, epa_w2=ew2, epa_normals=enr)


if __name__ == "__main__":
    import subprocess, tempfile
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        _one(sys.argv[2], sys.argv[3], sys.argv[4])
        sys.exit(0)
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.dirname(os.path.abspath(__file__))
    for prec, dt in (("f32", np.float32), ("f64", np.float64)):
        blob = {}
        for variant in ("", "_nofma"):
            for name, w in epa_workloads(dt):
                key = f"{name}{variant}"
                tmp = os.path.join(tempfile.mkdtemp(), "r.npz")
                rc = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", prec, variant, name, tmp]).returncode
                blob[name + "/sha"] = np.frombuffer(checksum(w).encode(), dtype=np.uint8)
                if rc != 0 or not os.path.exists(tmp):
                    print(prec, key, "reference run FAILED rc", rc)
                    blob[key + "/failed"] = np.array(True)
                    continue
                r = np.load(tmp)
                if int(r["cuda_error"]) != 0:
                    print(prec, key, "reference kernel left CUDA error", int(r["cuda_error"]))
                    blob[key + "/failed"] = np.array(True)
                    continue
                fields = ["dist", "w1", "w2", "repeatable"]
                if variant == "_nofma" and name == "boxes_overlap":  # the EPA stage alone, on the reference's own GJK results
                    fields += ["gjk_dist", "gjk_simplex", "epa_dist", "epa_w1", "epa_w2", "epa_normals"]
                for f in fields:
                    blob[f"{key}/{f}"] = r[f]
                print(prec, key, "penetrating", int((r["dist"] < 0).sum()), "of", r["dist"].shape[0], "repeatable", bool(r["repeatable"]))
        np.savez_compressed(os.path.join(out, f"epa_ref_cuda_{prec}.npz"), **blob)
