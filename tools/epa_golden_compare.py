import numpy as np, os, sys
sys.path.insert(0, os.getcwd())
from oracle.pyoracle import Oracle
from tests.golden.make_epa_golden import epa_workloads
for prec, dt in (("f32", np.float32), ("f64", np.float64)):
    g = np.load(f"gpurun_out/epa_ref_cuda_{prec}.npz")
    o = Oracle(prec)
    for name, w in epa_workloads(dt):
        r = o.gjk_epa_batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=8)
        for variant in ("", "_nofma"):
            if f"{name}{variant}/failed" in g.files:
                print(f"{prec} {name}{variant}: reference run failed (CUDA fault)"); continue
            d = g[f"{name}{variant}/dist"]; w1 = g[f"{name}{variant}/w1"]; w2 = g[f"{name}{variant}/w2"]
            pen = r["distances"] < 0
            same_sign = ((d < 0) == pen).mean()
            bit = (d.tobytes() == r["distances"].tobytes())
            biteq = (d.view(np.uint32 if prec=="f32" else np.uint64) == r["distances"].view(np.uint32 if prec=="f32" else np.uint64)).mean()
            tol = 1e-5 if prec == "f32" else 1e-10
            close = np.isclose(d, r["distances"], rtol=tol, atol=tol)
            close4 = np.isclose(d, r["distances"], rtol=1e-3, atol=1e-4)
            wclose = np.isclose(w1, r["witness1"], rtol=1e-3, atol=1e-3).all(1) & np.isclose(w2, r["witness2"], rtol=1e-3, atol=1e-3).all(1)
            print(f"{prec} {name}{variant}: n={d.size} pen={pen.mean():.3f} sign-agree={same_sign:.4f} bit-equal={biteq:.4f} within-tol({tol})={close.mean():.4f} within-1e-3={close4.mean():.4f} witnesses-1e-3={wclose.mean():.4f} penetrating-only-tol={close[pen].mean():.4f}")
