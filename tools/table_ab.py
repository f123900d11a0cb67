#!/usr/bin/env python
"""GPU-side A/B of library variants on the resident-store path (tools/build_variants.sh builds them).

    python tools/table_ab.py --workload hulls1k_f32 [--pairs N] --libs default,variants/libogjk_x.so:0,...

Each entry is <lib>[:table_lanes]; `default` = opengjk_b200/libopengjk_b200.so.  For every entry: build the
store, run the batch, compare distances + simplices byte for byte with the first entry's output, then time
`--steps` passes with CUDA events.  Prints one line per entry (M pairs/s, HBM-roofline fraction)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="hulls1k_f32")
    ap.add_argument("--pairs", type=int, default=0)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--libs", default="default")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--classes", default="")
    ap.add_argument("--out", default="")
    ap.add_argument("--accel-min", type=int, default=-1, help="cluster-table threshold (0 = no tables)")
    args = ap.parse_args()
    import torch
    import bench
    from opengjk_b200 import api, _lib
    desc, dt, maker, default_pairs = bench.WORKLOADS[args.workload][:4]
    n = args.pairs or default_pairs
    w = maker(n, args.seed)
    prec = "f64" if dt == np.float64 else "f32"
    sdt = api.SIMPLEX_F64 if prec == "f64" else api.SIMPLEX_F32
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_coords, d_off, d_cnt, d_pa, d_pb = t(w.coords.reshape(-1)), t(w.off), t(w.cnt), t(w.pa), t(w.pb)
    alg = w.algorithmic_bytes()
    peak = bench.peaks()[0]
    stream = torch.cuda.current_stream().cuda_stream
    base = None
    rows = []
    for ent in args.libs.split(","):
        lib, lanes, *rg = ent.split(":") if ":" in ent else (ent, "")  # lib[:table_lanes[:regroup1[:regroup2]]]
        path = _lib.LIB_PATH if lib == "default" else os.path.join(ROOT, "opengjk_b200", lib)
        eng = api.Engine(0, lib_path=path)
        if args.classes:
            eng.set_classes([tuple(int(x) for x in c.split(":")) for c in args.classes.split(",")])
        if lanes:
            eng.set_table_lanes(int(lanes))
        if rg:
            eng.set_regroup(int(rg[0]), int(rg[1]) if len(rg) > 1 else 0)
        if args.accel_min >= 0:
            eng.set_accel_min(args.accel_min)
        store = eng.store_create(d_coords, d_off, d_cnt, stream)
        d_dist = torch.zeros(n, dtype=d_coords.dtype, device=dev)
        d_simp = torch.zeros(n * sdt.itemsize, dtype=torch.uint8, device=dev)
        step = lambda: eng.store_gjk(store, d_pa, store, d_pb, n, d_simp, d_dist, stream)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        got = (d_dist.cpu().numpy().tobytes(), d_simp.cpu().numpy().tobytes())
        if base is None:
            base = got
        same = got == base
        best = 1e30
        tot = 0.0
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            best = min(best, ms)
            tot += ms
        row = {"lib": ent, "workload": args.workload, "pairs": n, "ms_best": best, "ms_mean": tot / 3,
               "mpairs_s": n / best / 1e3, "frac": alg / (best * 1e-3) / 1e9 / peak, "bytes_equal_to_first": same}
        rows.append(row)
        print(f"{ent:44s} {args.workload:14s} {row['mpairs_s']:9.1f} Mpairs/s  frac {row['frac']:.3f}  "
              f"{'same bytes' if same else 'BYTES DIFFER'}", flush=True)
        store.close()
        eng.close()
    if args.out:
        with open(args.out, "a") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
