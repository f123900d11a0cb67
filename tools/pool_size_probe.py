import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from opengjk_b200 import api, workloads as W
eng = api.Engine(0); dev = torch.device("cuda:0")
for npool in (512, 2048, 8192, 32768):
    w = W.large_hull_pool(200000, npool, 1000, 1000, 42, dtype=np.float32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    for amin in (0, 256):
        eng.set_accel_min(amin)
        store = eng.store_create(w.coords, w.off, w.cnt)
        pa, pb = t(w.pa), t(w.pb)
        simp = torch.empty(w.npairs * api.SIMPLEX_F32.itemsize, dtype=torch.uint8, device=dev)
        dist = torch.empty(w.npairs, dtype=torch.float32, device=dev)
        for _ in range(3): eng.store_gjk(store, pa, store, pb, w.npairs, simp, dist, 0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(5): eng.store_gjk(store, pa, store, pb, w.npairs, simp, dist, 0)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"pool {npool:6d} hulls ({npool*12/1024:.0f} MB) tables>={amin}: {w.npairs/ms/1e3:.1f} Mpairs/s", flush=True)
        store.close()
