"""Times ogjk_store_create on a device-resident pool (GPU): CUDA events around the call,
first call (pays module load and first allocations) and the best of three more.
    python tools/store_build_time.py hulls1k_f32 hulls1k10k_full_f32"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from opengjk_b200 import api
    dev = torch.device("cuda:0")
    eng = api.Engine(0)
    stream = torch.cuda.current_stream().cuda_stream
    for name in sys.argv[1:] or ["hulls1k_f32", "hulls1k10k_full_f32"]:
        desc, dt, maker, pairs = bench.WORKLOADS[name][:4]
        w = maker(1000, 42)  # the pool does not depend on the pair count
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        d_coords, d_off, d_cnt = t(w.coords.reshape(-1)), t(w.off), t(w.cnt)
        times = []
        for i in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            store = eng.store_create(d_coords, d_off, d_cnt, stream)
            torch.cuda.synchronize()
            times.append(1e3 * (time.perf_counter() - t0))
            nb, tb = store.nbytes, store.table_bytes
            store.close()
        print(json.dumps({"workload": name, "polytopes": int(w.cnt.size), "vertices": int(w.cnt.sum()),
                          "store_bytes": nb, "table_bytes": tb, "first_ms": times[0], "best_ms": min(times[1:]),
                          "all_ms": times}))


if __name__ == "__main__":
    main()
