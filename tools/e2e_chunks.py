#!/usr/bin/env python
"""GPU-side: e2e pairs/s of ogjk_gjk_host_* (two pinned pools) vs number of pipeline chunks."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS
from opengjk_b200 import api
eng = api.Engine(0)
for name in sys.argv[1:] or ["boxes_f32", "hulls64_f64"]:
    desc, dt, maker, pairs = WORKLOADS[name][:4]
    w = maker(pairs, 42)
    sdt = api.SIMPLEX_F64 if dt == np.float64 else api.SIMPLEX_F32
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    pools = [pin(x).numpy() for x in w.two_pools()]
    out = (pin(np.zeros(pairs, dtype=dt)).numpy(), pin(np.zeros(pairs * sdt.itemsize, dtype=np.uint8)).numpy().view(sdt))
    nbytes = sum(x.nbytes for x in pools)
    for chunks in (0, 2, 4, 8, 16):
        eng.set_pipeline_chunks(chunks)
        eng.gjk_host_pools(*pools, out=out)
        t0 = time.perf_counter()
        for _ in range(4):
            eng.gjk_host_pools(*pools, out=out)
        dt_s = (time.perf_counter() - t0) / 4
        print(f"{name:12s} chunks={chunks:2d}  {1e3*dt_s:7.2f} ms  {pairs/dt_s/1e6:8.1f} Mpairs/s  in {nbytes/1e6:.0f} MB ({nbytes/dt_s/1e9:.1f} GB/s) out {(out[0].nbytes+out[1].nbytes)/1e6:.0f} MB", flush=True)
