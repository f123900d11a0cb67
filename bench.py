#!/usr/bin/env python
"""bench.py -- GJK pair-queries/sec on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A "step" is one pass of the batched GJK hot path over one batch of synthetic
pairs.  Default workload = BASELINE.json configs[1]: 1 M random convex-hull
pairs, 8-64 vertices, fp64 (per GPU; weak scaling, each rank owns its own
contiguous slice of pairs, no data-path collective).

Printed JSON line (rank 0):
  value      whole-job pairs/s with inputs resident in HBM (CUDA events on the
             launching stream, barrier + synchronize on both sides, max over ranks)
  e2e        same metric through the host-buffer C-ABI call (ogjk_gjk_host_*):
             pinned host inputs -> H2D -> kernels -> D2H of distances + gkSimplex
  roofline   dominant kernel vs the measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline  the reference scalar library (oracle/_ref) or the oracle port on
             the box's host cores, bounded sample
`--impl reference` times the reference CPU implementation alone.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from opengjk_b200 import workloads as W  # noqa: E402

L2_BYTES = 126e6
# scheduler table: (hi_key, lanes per pair); key = (pad4(n1)+pad4(n2))/4; lanes 0 = register-resident kernel
CLASS_TABLE = [(4, 0), (32, 2), (128, 8), (2048, 32)]

WORKLOADS = {
    # name: (description, dtype, maker(npairs, seed), default pairs per GPU[, epa])
    "hulls64_f64": ("config2: random convex-hull pairs, 8-64 vertices, fp64", np.float64,
                    lambda n, seed: W.mixed_hulls(n, 8, 64, seed, np.float64), 1_000_000),
    "hulls64_f32": ("config2 in fp32: random convex-hull pairs, 8-64 vertices", np.float32,
                    lambda n, seed: W.mixed_hulls(n, 8, 64, seed, np.float32), 1_000_000),
    "boxes_f32": ("config3: oriented box-box (8-vertex) pairs, fp32", np.float32,
                  lambda n, seed: W.boxes(n, True, seed, dtype=np.float32), 8_000_000),
    "boxes_aligned_f32": ("config3a: axis-aligned unit cubes (tie stress), fp32", np.float32,
                          lambda n, seed: W.boxes(n, False, seed, dtype=np.float32), 8_000_000),
    "hulls1k_f32": ("config4a: indexed pairs of 1000-vertex hulls, fp32", np.float32,
                    lambda n, seed: W.large_hull_pool(n, 32768, 1000, 1000, seed, dtype=np.float32), 200_000),
    "hulls1k10k_f32": ("config4: indexed pairs of 1k-10k-vertex hulls, fp32", np.float32,
                       lambda n, seed: W.large_hull_pool(n, 2048, 1000, 10000, seed, dtype=np.float32), 50_000),
    "hulls1k10k_full_f32": ("config4 at full size: 1 M indexed pairs over a pool of 8192 hulls with 1k-10k vertices (540 MB), fp32",
                            np.float32, lambda n, seed: W.large_hull_pool(n, 8192, 1000, 10000, seed, dtype=np.float32), 1_000_000),
    "boxes100m_f32": ("config3 at full size: 100 M oriented box-box pairs generated on the device, fp32", np.float32,
                      None, 100_000_000),
    "epa_f32": ("config5: intersecting 8-64-vertex hull pairs, GJK + EPA, fp32", np.float32,
                lambda n, seed: W.random_hulls(n, 8, 64, seed, 1.0, np.float32, 0.5), 1_000_000, True),
}


def traffic_for(name):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    try:
        return json.load(open(p)).get(name)
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the measured region (NVML every 20 ms;
    falls back to polling nvidia-smi)."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.sm, self.bits, self.max_mhz, self._stop = index, [], 0, None, threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            while not self._stop.is_set():
                self.sm.append(int(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                try:
                    self.bits |= int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    self.bits |= int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self._stop.wait(0.02)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                r = [x.strip() for x in out.split(",")]
                self.sm.append(int(r[0]))
                self.max_mhz = int(r[1])
                for i, n in enumerate(names):
                    if r[2 + i].lower().startswith("active"):
                        self.bits |= self.REASONS[n]
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"], "samples": 0}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz,
                "reasons": [n for n, b in self.REASONS.items() if self.bits & b], "samples": len(sm)}


def cpu_baseline(w, prec, budget_s=10.0):
    """Reference scalar library (or the oracle port) on all host cores, bounded sample."""
    from oracle.pyoracle import Oracle, Reference
    cores = os.cpu_count() or 1
    n = min(w.npairs, 100_000)
    if Reference.available(prec):
        kind, ref = "reference", Reference(prec)
        prep = Reference.Prepared(ref, w.coords, w.off, w.cnt, w.pa[:n], w.pb[:n])
        run = lambda: prep.run(cores)
    else:
        kind, o = "port", Oracle(prec)
        run = lambda: o.batch(w.coords, w.off, w.cnt, w.pa[:n], w.pb[:n], nthreads=cores)
    run()  # warm
    t0 = time.perf_counter()
    reps = 0
    while True:
        run()
        reps += 1
        if time.perf_counter() - t0 > budget_s or reps >= 20000:
            break
    dt = time.perf_counter() - t0
    return {"value": n * reps / dt, "unit": "pairs/s", "cores": cores, "kind": kind,
            "sample": f"first {n} pairs of the workload x {reps} passes ({dt:.1f} s), OpenMP split over pairs"}


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation, rank 0 only."""
    if rank != 0:
        return
    desc, dt, maker, default_pairs = WORKLOADS[args.workload][:4]
    prec = "f64" if dt == np.float64 else "f32"
    from oracle.pyoracle import Oracle, Reference
    cores = os.cpu_count() or 1
    n = min(args.pairs or default_pairs, args.ref_sample)
    maker = maker or (lambda n_, seed_: W.boxes(n_, True, seed_, dtype=np.float32))  # device-generated workloads
    w = maker(n, args.seed)
    if Reference.available(prec):
        kind, ref = "reference", Reference(prec)
        prep = Reference.Prepared(ref, w.coords, w.off, w.cnt, w.pa, w.pb)
        run = lambda: prep.run(cores)
    else:
        kind, o = "port", Oracle(prec)
        run = lambda: o.batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=cores)
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt_s = time.perf_counter() - t0
    v = n * args.steps / dt_s
    line = {
        "impl": "reference", "metric": "gjk_pair_queries_per_sec", "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt_s / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": prec, "data": "synthetic",
        "config": {"workload": desc, "pairs_per_step": n, "note": "CPU reference, bounded sample of the workload"},
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": kind,
                         "sample": f"{n} pairs per step, all host threads (OpenMP static split over pairs)"},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="hulls64_f64", choices=sorted(WORKLOADS))
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU (default: the workload's full size)")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-sample", type=int, default=200_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extra", default="", help="comma list of further workloads reported under other_workloads")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from opengjk_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng = api.Engine(local_rank)
    if os.environ.get("OGJK_CLASSES"):  # e.g. "4:0,24:4,64:8,192:16,2048:32"
        CLASS_TABLE[:] = [tuple(int(x) for x in c.split(":")) for c in os.environ["OGJK_CLASSES"].split(",")]
    eng.set_classes(CLASS_TABLE)
    if os.environ.get("OGJK_ACCEL_MIN"):  # tuning hook: 0 = no cluster tables in the resident store
        eng.set_accel_min(int(os.environ["OGJK_ACCEL_MIN"]))
    if os.environ.get("OGJK_PASSES"):  # e.g. "2:4:3" = passes:iters1:iters2
        eng.set_passes(*[int(x) for x in os.environ["OGJK_PASSES"].split(":")])
    if os.environ.get("OGJK_TABLE_LANES"):
        eng.set_table_lanes(int(os.environ["OGJK_TABLE_LANES"]))
    hbm_peak, peak_src = peaks()

    def measure(name, pairs, steps, warmup, with_cpu):
        desc, dt, maker, default_pairs = WORKLOADS[name][:4]
        epa = len(WORKLOADS[name]) > 4 and WORKLOADS[name][4]
        prec = "f64" if dt == np.float64 else "f32"
        n = pairs or default_pairs
        w = maker(n, args.seed + rank)  # each rank owns its own slice of the job
        sdt = api.SIMPLEX_F64 if prec == "f64" else api.SIMPLEX_F32

        # ---- pinned host copies (e2e inputs) and device-resident copies (value) ----
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        h_coords, h_off, h_cnt, h_pa, h_pb = pin(w.coords.reshape(-1)), pin(w.off), pin(w.cnt), pin(w.pa), pin(w.pb)
        h_dist = torch.empty(n, dtype=h_coords.dtype).pin_memory()
        h_simp = torch.empty(n * sdt.itemsize, dtype=torch.uint8).pin_memory()
        d_coords, d_off, d_cnt, d_pa, d_pb = (t.to(dev) for t in (h_coords, h_off, h_cnt, h_pa, h_pb))
        d_dist = torch.empty(n, dtype=h_coords.dtype, device=dev)
        d_simp = torch.empty(n * sdt.itemsize, dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream().cuda_stream
        in_bytes = sum(t.numel() * t.element_size() for t in (h_coords, h_off, h_cnt, h_pa, h_pb))
        out_bytes = h_dist.numel() * h_dist.element_size() + h_simp.numel()

        # The engine's device-resident format is the packed polytope store (built once,
        # outside the timed region -- the counterpart of the reference's
        # allocate_and_copy_device_arrays); a step is one ogjk_store_gjk_* call over it.
        t0 = time.perf_counter()
        store = eng.store_create(d_coords, d_off, d_cnt, stream)
        torch.cuda.synchronize()
        store_build_ms = 1e3 * (time.perf_counter() - t0)

        if epa:
            d_w1, d_w2, d_nr = (torch.empty(n * 3, dtype=h_coords.dtype, device=dev) for _ in range(3))

        def step():
            if epa:
                eng.store_gjk_epa(store, d_pa, store, d_pb, n, d_simp, d_dist, d_w1, d_w2, d_nr, stream)
            else:
                eng.store_gjk(store, d_pa, store, d_pb, n, d_simp, d_dist, stream)

        def step_from_raw():
            eng.gjk_device(prec, n, d_coords, d_off, d_cnt, d_pa, d_coords, d_off, d_cnt, d_pb, d_simp, d_dist,
                           api.MODE_AUTO, stream)

        # ---- device-resident timing -------------------------------------------------
        eng.set_timing(False)
        for _ in range(warmup):
            step()
        barrier()
        launches0 = eng.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as clk:
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            launches = eng.launch_count - launches0  # kernels launched inside the timed region
            barrier()
            # keep the same load on the GPU for >= 0.3 s so the clock record has samples
            # taken under load even when K steps last only a few milliseconds (untimed)
            t_load = time.perf_counter()
            while time.perf_counter() - t_load < 0.3 or (len(clk.sm) < 8 and time.perf_counter() - t_load < 3.0):
                step()
                torch.cuda.synchronize()
                time.sleep(0)  # let the sampler thread run
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
        value = n * world * steps / (ms_max * 1e-3)

        # same job starting from the raw xyz-interleaved arrays (repack inside the step)
        raw_ms = float("nan")
        if not epa:
            for _ in range(2):
                step_from_raw()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(max(2, steps // 2)):
                step_from_raw()
            e1.record()
            torch.cuda.synchronize()
            raw_ms = e0.elapsed_time(e1) / max(2, steps // 2)

        # ---- per-phase durations (events around each launch, same stream) -----------
        eng.set_timing(True)
        sort_ms, class_ms = [], []
        for _ in range(max(3, min(steps, 10))):
            step()
            torch.cuda.synchronize()
            lt = eng.last_timing()
            sort_ms.append(lt["sort"])
            class_ms.append(lt["classes"])  # last slot = EPA when it ran
        eng.set_timing(False)
        class_avg = [float(x) for x in np.mean(np.array(class_ms), axis=0)]
        dom = int(np.argmax(class_avg))
        alg_bytes = w.algorithmic_bytes()
        # the dominant kernel is charged the whole batch's algorithmic bytes (conservative:
        # other size classes, if any, moved some of them)
        achieved = alg_bytes / (class_avg[dom] * 1e-3) / 1e9
        tname = "double" if prec == "f64" else "float"
        if epa and dom == 5:
            kname = f"epa_warp_kernel<{tname}>"
        else:
            lanes = CLASS_TABLE[dom][1]
            amin = int(os.environ.get("OGJK_ACCEL_MIN", 512))
            # as run_store_batch routes: pairs whose two bodies both carry a table run the table kernel,
            # timed with the last class
            if store.table_bytes > 0 and dom == len(CLASS_TABLE) - 1 and int(w.cnt.min()) >= amin:
                kname = f"gjk_table_kernel<{tname},{int(os.environ.get('OGJK_TABLE_LANES', 32))}>"
            else:
                kname = f"gjk_small_kernel<{tname}>" if lanes == 0 else f"gjk_group_kernel<{tname},{lanes}>"
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": traffic_for(name), "kernel": kname, "kernel_ms": class_avg[dom],
                    "step_phases_ms": {"sort": float(np.mean(sort_ms)), "classes": class_avg},
                    "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src}

        # ---- end to end through the host-buffer C-ABI call --------------------------
        o = (h_dist.numpy(), h_simp.numpy().view(sdt))
        hc, ho, hk, ha, hb = (t.numpy() for t in (h_coords, h_off, h_cnt, h_pa, h_pb))
        e2e_steps = max(2, min(steps, 5))
        pair_owned = bool(np.array_equal(w.pa, np.arange(0, 2 * n, 2)) and np.array_equal(w.pb, w.pa + 1))
        if pair_owned and not epa:
            # the reference's bd1[] / bd2[] call shape: two pinned pools, pair p = polytope p of each
            # (chunk-pipelined H2D / kernels / D2H inside the library)
            pools = [pin(x).numpy() for x in w.two_pools()]
            e2e_call = lambda: eng.gjk_host_pools(*pools, out=o)
            in_bytes = sum(x.nbytes for x in pools)
        elif epa:
            e2e_call = lambda: eng.gjk_epa_host(hc, ho, hk, ha, hb)
        else:
            e2e_call = lambda: eng.gjk_host(hc, ho, hk, ha, hb, True, o)
        e2e_call()  # warm (allocates staging)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_call()
        barrier()
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": n * world * e2e_steps / float(t.item()), "unit": "pairs/s", "h2d_bytes_per_step": in_bytes,
               "d2h_bytes_per_step": out_bytes + (n * 9 * h_dist.element_size() if epa else 0), "steps": e2e_steps,
               "api": (f"ogjk_gjk_epa_host_{prec}" if epa else f"ogjk_gjk_host_{prec}")
               + " (pinned host buffers in, distances + gkSimplex" + (" + witnesses + normals" if epa else "") + " out)"}

        # quick self-check on a sample against the oracle (checker only, outside any timed region)
        res = {"workload": desc, "value": value, "from_raw_arrays": (None if epa else {"value": n * world / (raw_ms * 1e-3), "ms_per_step": raw_ms,
                                                                                   "note": "repack into the store inside the step"}),
               "store_build_ms": store_build_ms, "store_bytes": store.nbytes, "ms_per_step": ms_max / steps, "launches": launches, "roofline": roofline,
               "e2e": e2e, "clocks": clk.summary(), "prec": prec, "pairs_per_gpu": n,
               "input_bytes_per_gpu": in_bytes}
        if with_cpu and rank == 0 and epa:
            from oracle.pyoracle import Oracle
            o_ = Oracle(prec)
            m = min(n, 20000)
            cores = os.cpu_count() or 1
            t0 = time.perf_counter()
            ref = o_.gjk_epa_batch(w.coords, w.off, w.cnt, w.pa[:m], w.pb[:m], nthreads=cores)
            dt_s = time.perf_counter() - t0
            res["cpu_baseline"] = {"value": m / dt_s, "unit": "pairs/s", "cores": cores, "kind": "port",
                                   "sample": f"first {m} pairs, GJK + EPA restatement (the reference has no CPU EPA)"}
            res["sample_parity"] = {"pairs": m, "distance_bits_equal": bool(
                d_dist[:m].cpu().numpy().tobytes() == ref["distances"].tobytes()),
                "witness_bits_equal": bool(d_w1[:3 * m].cpu().numpy().tobytes() == ref["witness1"].tobytes())}
        elif with_cpu and rank == 0:
            res["cpu_baseline"] = cpu_baseline(w, prec)
            from oracle.pyoracle import Oracle
            m = min(n, 20000)
            dref, sref = Oracle(prec).batch(w.coords, w.off, w.cnt, w.pa[:m], w.pb[:m], nthreads=os.cpu_count() or 1)
            got_d = d_dist[:m].cpu().numpy()
            res["sample_parity"] = {"pairs": m, "distance_bits_equal": bool(got_d.tobytes() == dref.tobytes()),
                                    "simplex_bytes_equal": bool(
                                        d_simp[:m * sdt.itemsize].cpu().numpy().tobytes() == sref.tobytes())}
        store.close()
        del d_coords, d_simp, d_dist
        torch.cuda.empty_cache()
        return res

    def device_boxes(n, seed):
        """Config #3 inputs generated on the GPU in 10 M-pair slabs (same construction as
        workloads.boxes: half-extents U[0.5,1.5], uniform random rotation, body 2 offset 6(u-0.5))."""
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        coords = torch.empty(2 * n * 8 * 3, dtype=torch.float32, device=dev)
        box = torch.tensor(W._BOX, dtype=torch.float32, device=dev)
        slab = 10_000_000
        for lo in range(0, n, slab):
            m = min(slab, n - lo)
            nb = 2 * m
            half = 0.5 + torch.rand((nb, 1, 3), generator=g, device=dev)
            q = torch.randn((nb, 4), generator=g, device=dev)
            q = q / q.norm(dim=1, keepdim=True)
            qw, qx, qy, qz = q.unbind(1)
            rot = torch.stack([1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw),
                               2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw),
                               2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)], 1).view(nb, 3, 3)
            pts = torch.einsum("bij,bvj->bvi", rot, box[None] * half)
            shift = torch.zeros((nb, 1, 3), device=dev)
            shift[1::2, 0] = 6.0 * (torch.rand((m, 3), generator=g, device=dev) - 0.5)
            coords[2 * lo * 24:2 * (lo + m) * 24] = (pts + shift).reshape(-1)
            del half, q, rot, pts, shift
        cnt = torch.full((2 * n,), 8, dtype=torch.int32, device=dev)
        off = torch.arange(2 * n, dtype=torch.int64, device=dev) * 8
        pa = torch.arange(0, 2 * n, 2, dtype=torch.int32, device=dev)
        return coords, off, cnt, pa, pa + 1

    def measure_device_boxes(name, pairs, steps, warmup, with_cpu):
        """100 M-pair box batch: inputs never exist on the host; parity is checked on samples
        copied back (first 20 k pairs and a strided 20 k), e2e on an 8 M-pair sub-batch."""
        desc = WORKLOADS[name][0]
        n = pairs or WORKLOADS[name][3]
        prec, sdt = "f32", api.SIMPLEX_F32
        d_coords, d_off, d_cnt, d_pa, d_pb = device_boxes(n, args.seed + rank)
        d_dist = torch.empty(n, dtype=torch.float32, device=dev)
        d_simp = torch.empty(n * sdt.itemsize, dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream().cuda_stream
        t0 = time.perf_counter()
        store = eng.store_create(d_coords, d_off, d_cnt, stream)
        torch.cuda.synchronize()
        store_build_ms = 1e3 * (time.perf_counter() - t0)
        step = lambda: eng.store_gjk(store, d_pa, store, d_pb, n, d_simp, d_dist, stream)
        eng.set_timing(False)
        for _ in range(warmup):
            step()
        barrier()
        launches0 = eng.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as clk:
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            launches = eng.launch_count - launches0
            barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
        eng.set_timing(True)
        cls = []
        for _ in range(3):
            step()
            torch.cuda.synchronize()
            lt = eng.last_timing()
            cls.append([lt["sort"]] + lt["classes"])
        eng.set_timing(False)
        avg = [float(x) for x in np.mean(np.array(cls), axis=0)]
        alg_bytes = n * (16 * 12 + 108 + 4)
        kms = avg[1]  # class 0 = register-resident kernel
        roofline = {"bound": "hbm", "achieved": alg_bytes / (kms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": alg_bytes / (kms * 1e-3) / 1e9 / hbm_peak, "traffic": None, "kernel": "gjk_small_kernel<float>",
                    "kernel_ms": kms, "step_phases_ms": {"sort": avg[0], "classes": avg[1:]},
                    "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src}
        # samples back to the host: parity vs the oracle, CPU baseline, e2e sub-batch
        def host_sample(idx):
            idx = np.asarray(idx)
            c = torch.cat([d_coords.view(-1, 48)[idx], ]).cpu().numpy().reshape(-1, 3)  # both boxes of a pair are adjacent
            m = len(idx)
            cnt = np.full(2 * m, 8, dtype=np.int32)
            off = np.arange(2 * m, dtype=np.int64) * 8
            pa = np.arange(0, 2 * m, 2, dtype=np.int32)
            return W.Workload("boxes_sample", c, off, cnt, pa, pa + 1)
        res = {"workload": desc, "value": n * world * steps / (ms_max * 1e-3), "ms_per_step": ms_max / steps,
               "launches": launches, "roofline": roofline, "clocks": clk.summary(), "prec": prec, "pairs_per_gpu": n,
               "input_bytes_per_gpu": int(d_coords.numel() * 4 + d_off.numel() * 8 + d_cnt.numel() * 4 + 2 * d_pa.numel() * 4),
               "from_raw_arrays": None, "store_build_ms": store_build_ms, "store_bytes": store.nbytes}
        if rank == 0:
            from oracle.pyoracle import Oracle
            checks = {}
            for label, idx in (("first", np.arange(20000)), ("strided", np.arange(0, n, max(1, n // 20000))[:20000])):
                ws = host_sample(idx)
                dref, sref = Oracle(prec).batch(ws.coords, ws.off, ws.cnt, ws.pa, ws.pb, nthreads=os.cpu_count() or 1)
                ti = torch.from_numpy(idx).to(dev)
                got_d = d_dist[ti].cpu().numpy()
                got_s = d_simp.view(-1, sdt.itemsize)[ti].cpu().numpy().reshape(-1).view(sdt)
                checks[label] = {"pairs": len(idx), "distance_bits_equal": bool(got_d.tobytes() == dref.tobytes()),
                                 "simplex_bytes_equal": bool(got_s.tobytes() == sref.tobytes())}
            res["sample_parity"] = checks
            if with_cpu:
                res["cpu_baseline"] = cpu_baseline(host_sample(np.arange(100000)), prec)
        # e2e on an 8 M-pair sub-batch through the host-buffer API (two pinned pools)
        m = min(n, 8_000_000)
        sub = d_coords[: m * 48].view(m, 2, 24)
        pin = lambda a: a.contiguous().cpu().pin_memory().numpy()
        c1, c2 = pin(sub[:, 0].reshape(-1)), pin(sub[:, 1].reshape(-1))
        k = torch.full((m,), 8, dtype=torch.int32).pin_memory().numpy()
        o = (torch.arange(m, dtype=torch.int64) * 8).pin_memory().numpy()
        out = (torch.empty(m, dtype=torch.float32).pin_memory().numpy(),
               torch.empty(m * sdt.itemsize, dtype=torch.uint8).pin_memory().numpy().view(sdt))
        eng.gjk_host_pools(c1, o, k, c2, o, k, out=out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            eng.gjk_host_pools(c1, o, k, c2, o, k, out=out)
        barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res["e2e"] = {"value": m * world * 3 / float(t.item()), "unit": "pairs/s",
                      "h2d_bytes_per_step": int(c1.nbytes + c2.nbytes + 2 * (k.nbytes + o.nbytes)),
                      "d2h_bytes_per_step": int(out[0].nbytes + out[1].nbytes), "steps": 3,
                      "api": "ogjk_gjk_host_f32 on an 8 M-pair sub-batch (two pinned pools, chunk-pipelined)"}
        store.close()
        del d_coords, d_simp, d_dist
        torch.cuda.empty_cache()
        return res

    _measure = measure

    def measure(name, pairs, steps, warmup, with_cpu):
        if WORKLOADS[name][2] is None:
            return measure_device_boxes(name, pairs, steps, warmup, with_cpu)
        return _measure(name, pairs, steps, warmup, with_cpu)

    main_res = measure(args.workload, args.pairs, args.steps, args.warmup, not args.no_cpu_baseline)
    others = {}
    for name in [x for x in args.extra.split(",") if x]:
        r = measure(name, 0, max(3, args.steps // 2), 3, not args.no_cpu_baseline)
        others[name] = {k: r[k] for k in ("workload", "value", "ms_per_step", "roofline", "e2e", "prec", "pairs_per_gpu", "from_raw_arrays")
                        if k in r}
        if "cpu_baseline" in r:
            others[name]["cpu_baseline"] = r["cpu_baseline"]
            others[name]["sample_parity"] = r.get("sample_parity")

    if rank == 0:
        line = {
            "metric": "gjk_pair_queries_per_sec", "value": main_res["value"], "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": main_res["prec"], "data": "synthetic",
            "config": {"workload": main_res["workload"], "pairs_per_gpu": main_res["pairs_per_gpu"],
                       "sharding": "contiguous pair slice per rank, no data-path collective",
                       "l2": (f"inputs {main_res['input_bytes_per_gpu'] / 1e6:.0f} MB per GPU "
                              + ("> 126 MB L2 (no flush needed)" if main_res["input_bytes_per_gpu"] > L2_BYTES
                                 else "< L2: numbers are L2-warm")),
                       "resident_format": "packed polytope store (ogjk_store_create, outside the timed region)",
                       "scheduler_classes": [list(c) for c in CLASS_TABLE]},
            "roofline": main_res["roofline"], "e2e": main_res["e2e"], "gpu_launches": main_res["launches"],
            "from_raw_arrays": main_res["from_raw_arrays"], "store_build_ms": main_res["store_build_ms"],
            "clocks": main_res["clocks"],
        }
        if "cpu_baseline" in main_res:
            line["cpu_baseline"] = main_res["cpu_baseline"]
            line["sample_parity"] = main_res["sample_parity"]
        if others:
            line["other_workloads"] = others
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
