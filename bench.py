#!/usr/bin/env python
"""bench.py -- GJK pair-queries/sec on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--extra a,b|none]
                    [--impl reference]

A "step" is one pass of the batched GJK hot path over one batch of synthetic pairs.
Main workload = BASELINE.json configs[1]: 1 M random convex-hull pairs, 8-64 vertices,
fp64, per GPU (weak scaling: every rank owns its own contiguous slice of the job, no
data-path collective).  By default the other north-star configurations are measured in
the same run and reported under `other_workloads`:
  boxes100m_f32        config 3: ONE 100 M-pair oriented-box batch split over the N GPUs (strong)
  hulls1k_f32          config 4a: 1 M indexed pairs over 32768 1000-vertex hulls (the ">= 70 % of HBM" target)
  hulls1k10k_full_f32  config 4: 1 M indexed pairs over 8192 hulls with 1 k-10 k vertices
  epa10m_f32           config 5: 10 M intersecting pairs, GJK + EPA

Printed JSON line (rank 0):
  value      whole-job pairs/s with inputs resident in HBM (CUDA events on the launching
             stream, barrier + synchronize on both sides, max over ranks)
  e2e        same metric through the host-buffer C-ABI call: pinned host inputs -> H2D ->
             kernels -> D2H of distances + gkSimplex, with a bare-copy probe beside it
  roofline   dominant kernel vs the measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline  the unmodified reference scalar library (oracle/_ref) on the box's host
             cores (all cores and one thread), bounded strided sample
`--impl reference` times the reference CPU implementation alone on the same workload.
"""
import argparse
import contextlib
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
DEFAULT_EXTRAS = ["boxes100m_f32", "hulls1k_f32", "hulls1k10k_full_f32", "epa10m_f32"]

WORKLOADS = {
    # name: (description, dtype, maker(npairs, seed) or None for device-generated, default pairs[, epa])
    "hulls64_f64": ("config2: random convex-hull pairs, 8-64 vertices, fp64", np.float64,
                    lambda n, seed: W.mixed_hulls(n, 8, 64, seed, np.float64), 1_000_000),
    "hulls64_f32": ("config2 in fp32: random convex-hull pairs, 8-64 vertices", np.float32,
                    lambda n, seed: W.mixed_hulls(n, 8, 64, seed, np.float32), 1_000_000),
    "boxes_f32": ("config3 (8 M sample): oriented box-box (8-vertex) pairs, fp32", np.float32,
                  lambda n, seed: W.boxes(n, True, seed, dtype=np.float32), 8_000_000),
    "boxes_aligned_f32": ("config3a: axis-aligned unit cubes (tie stress), fp32", np.float32,
                          lambda n, seed: W.boxes(n, False, seed, dtype=np.float32), 8_000_000),
    "hulls1k_f32": ("config4a: 1 M indexed pairs over a pool of 32768 1000-vertex hulls (393 MB), fp32", np.float32,
                    lambda n, seed: W.large_hull_pool(n, 32768, 1000, 1000, seed, dtype=np.float32), 1_000_000),
    "hulls1k_f64": ("config4a in fp64: indexed pairs of 1000-vertex hulls", np.float64,
                    lambda n, seed: W.large_hull_pool(n, 16384, 1000, 1000, seed, dtype=np.float64), 500_000),
    "hulls1k10k_f32": ("config4 (small pool): indexed pairs of 1k-10k-vertex hulls, fp32", np.float32,
                       lambda n, seed: W.large_hull_pool(n, 2048, 1000, 10000, seed, dtype=np.float32), 50_000),
    "hulls1k10k_full_f32": ("config4: 1 M indexed pairs over a pool of 8192 hulls with 1k-10k vertices (540 MB), fp32",
                            np.float32, lambda n, seed: W.large_hull_pool(n, 8192, 1000, 10000, seed, dtype=np.float32), 1_000_000),
    "boxes100m_f32": ("config3: 100 M oriented box-box pairs generated on the device, fp32", np.float32,
                      None, 100_000_000),
    "epa_f32": ("config5 (1 M sample): intersecting 8-64-vertex hull pairs, GJK + EPA, fp32", np.float32,
                lambda n, seed: W.random_hulls(n, 8, 64, seed, 1.0, np.float32, 0.5), 1_000_000, True),
    "epa10m_f32": ("config5: 10 M intersecting 8-64-vertex hull pairs generated on the device, GJK + EPA, fp32", np.float32,
                   None, 10_000_000, True),
}
STRONG = {"boxes100m_f32"}  # ONE batch of the stated size split over the ranks (everything else: per-rank size)

# What limits each kernel, with the figures of the committed ncu capture (profiles/): the
# HBM fraction is always reported, but only the large-hull kernels are HBM-shaped.
NCU_NOTES = {
    "hulls64_f64": {"limiter": "scan load latency (long_scoreboard 4.7 cycles per issue) at 16 warps per SM / FP64 pipe / "
                               "warp divergence: 63 % of a warp's lane-iterations are live, 0.65 of those lanes converge "
                               "(not HBM-bound)",
                    "lanes_active": 13.1, "issue_active_pct": 42.8, "pipe_pct": {"fp64": 31.6, "alu": 23.3},
                    "source": "profiles/r2_hulls64_f64_ncu_summary.txt"},
    "hulls1k_f32": {"limiter": "instruction issue (3.5 k warp instructions per pair, 61 % issue-active at 32 warps per SM), then "
                               "the dependent cluster-record fetches (long_scoreboard); DRAM traffic is below the algorithmic bytes",
                    "lanes_active": 27.5, "issue_active_pct": 60.7, "source": "profiles/r2_hulls1k_f32_ncu_summary.txt"},
    "hulls1k10k_full_f32": {"limiter": "instruction issue (64 % issue-active at 28 warps per SM); tile and cluster pruning read "
                                       "a sixth of the algorithmic bytes, hence a roofline fraction above 1",
                            "issue_active_pct": 63.8, "source": "profiles/r2_hulls1k10k_full_f32_ncu_summary.txt"},
    "boxes100m_f32": {"limiter": "issue slots / warp divergence (not HBM-bound)", "lanes_active": 11.3,
                      "issue_active_pct": 68.1, "pipe_pct": {"alu": 52.5, "fma": 33.7},
                      "source": "profiles/r2_boxes_f32_ncu_summary.txt (same kernel, 8 M-pair batch)"},
    "epa10m_f32": {"limiter": "instruction issue (84 % issue-active; LSU pipe 51 %: the polytope lives in shared memory); "
                              "mean 16.5 expansions per pair (not HBM-bound)", "lanes_active": 21.0,
                   "issue_active_pct": 84.2, "pipe_pct": {"alu": 57.1, "fma": 24.4, "lsu": 51.2},
                   "source": "profiles/r2_epa_f32_ncu_summary.txt (same kernel, 1 M-pair batch)"},
}


def traffic_for(name):
    """DRAM bytes per launch of the dominant kernel from the committed ncu captures, with the file it came from."""
    for f in ("r2_traffic.json", "r1_traffic.json"):
        p = os.path.join(ROOT, "profiles", f)
        try:
            v = json.load(open(p)).get(name)
        except Exception:
            v = None
        if isinstance(v, (int, float)):
            return v, f"profiles/{f} (ncu --set full capture of the same kernel, committed; not measured in this run)"
    return None, None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


@contextlib.contextmanager
def quiet_stdout():
    """The reference prints a banner on its iteration-cap path; keep it out of our stdout
    (and its terminal cost out of the CPU timing) while reference code runs."""
    sys.stdout.flush()
    saved = os.dup(1)
    null = os.open(os.devnull, os.O_WRONLY)
    os.dup2(null, 1)
    try:
        yield
    finally:
        os.dup2(saved, 1)
        os.close(null)
        os.close(saved)


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


ORIG_AFFINITY = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None


@contextlib.contextmanager
def all_host_cores():
    """CPU baselines and the one-process multi-GPU run use every core the job was given, not just the
    NUMA node a rank was bound to."""
    cur = os.sched_getaffinity(0) if ORIG_AFFINITY is not None else None
    if cur is not None and cur != ORIG_AFFINITY:
        os.sched_setaffinity(0, ORIG_AFFINITY)
    try:
        yield
    finally:
        if cur is not None and cur != ORIG_AFFINITY:
            os.sched_setaffinity(0, cur)


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off, BEFORE any pinned host
    buffer is allocated (first touch then places the pools next to the GPU's PCIe root).  Round 1
    measured the host-buffer path at 0.48 / 0.37 of linear on 4 / 8 GPUs with unbound ranks.
    Best effort: returns a short description or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:      # nvml prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return f"numa node {node} ({len(cpus)} cpus)"
    except Exception:
        return None


def strided_sample(w, m):
    """m pairs of the workload taken at a fixed stride (keeps the mix of a workload built from slices)."""
    if w.npairs <= m:
        return w, f"all {w.npairs} pairs"
    st = w.npairs // m
    return W.Workload(w.name, w.coords, w.off, w.cnt, w.pa[::st][:m].copy(), w.pb[::st][:m].copy()), \
        f"{m} pairs at stride {st} over the workload"


def _timed_passes(run, n, budget_s, max_reps=20000):
    run()  # warm
    t0 = time.perf_counter()
    reps = 0
    while True:
        run()
        reps += 1
        if time.perf_counter() - t0 > budget_s or reps >= max_reps:
            break
    dt = time.perf_counter() - t0
    return n * reps / dt, reps, dt


def cpu_baseline(w, prec, budget_s=10.0, epa=False, sample_note=None):
    """The reference's scalar code (oracle/_ref; the oracle port when it is absent or for EPA,
    which the reference has only in CUDA) on the host cores: all cores and one thread."""
    from oracle.pyoracle import Oracle, Reference
    cores = os.cpu_count() or 1
    ws, note = strided_sample(w, 20_000 if epa else 100_000)
    note = sample_note or note
    n = ws.npairs
    if epa:
        kind, o = "port", Oracle(prec)
        run = lambda nt: (lambda: o.gjk_epa_batch(ws.coords, ws.off, ws.cnt, ws.pa, ws.pb, nthreads=nt))
    elif Reference.available(prec):
        kind, ref = "reference", Reference(prec)
        prep = Reference.Prepared(ref, ws.coords, ws.off, ws.cnt, ws.pa, ws.pb)
        run = lambda nt: (lambda: prep.run(nt))
    else:
        kind, o = "port", Oracle(prec)
        run = lambda nt: (lambda: o.batch(ws.coords, ws.off, ws.cnt, ws.pa, ws.pb, nthreads=nt))
    with quiet_stdout(), all_host_cores():
        v_all, reps, dt = _timed_passes(run(cores), n, budget_s)
        # one thread: a 1/8 slice of the sample, a quarter of the budget
        m1 = max(1, n // 8)
        w1 = W.Workload(ws.name, ws.coords, ws.off, ws.cnt, ws.pa[:m1 * 8:8].copy(), ws.pb[:m1 * 8:8].copy())
        if epa:
            run1 = lambda: o.gjk_epa_batch(w1.coords, w1.off, w1.cnt, w1.pa, w1.pb, nthreads=1)
        elif kind == "reference":
            prep1 = Reference.Prepared(ref, w1.coords, w1.off, w1.cnt, w1.pa, w1.pb)
            run1 = lambda: prep1.run(1)
        else:
            run1 = lambda: o.batch(w1.coords, w1.off, w1.cnt, w1.pa, w1.pb, nthreads=1)
        v_one, reps1, dt1 = _timed_passes(run1, w1.npairs, max(1.0, budget_s / 4))
    ref_cuda = reference_cuda_baseline(ws, epa)
    return {"value": v_all, "unit": "pairs/s", "cores": cores, "kind": kind, "reference_cuda_kernels": ref_cuda,
            "sample": f"{note} x {reps} passes ({dt:.1f} s), OpenMP split over pairs"
                      + (" (GJK + EPA restatement: the reference has no CPU EPA)" if epa else ""),
            "single_thread": {"value": v_one, "unit": "pairs/s", "cores": 1,
                              "sample": f"{w1.npairs} pairs (every 8th of the sample) x {reps1} passes ({dt1:.1f} s)"},
            "simd_baseline": "unavailable (Google Highway is not present offline; simd/CMakeLists.txt:58-75 fetches it)"}


def reference_cuda_baseline(ws, epa):
    """The reference's own CUDA kernels (gpu/opengjk_gpu.cu built unmodified into oracle/_ref), kernel-only through its
    mid-level API, on this GPU and on a bounded sample of the same pairs: its allocate_and_copy_device_arrays copies
    every pair's two bodies separately, so the sample is cut to <= 256 MB of vertices.  Runs in a child process."""
    import subprocess
    import tempfile
    per_pair = float(ws.cnt[ws.pa].astype(np.int64).sum() + ws.cnt[ws.pb].astype(np.int64).sum()) / ws.npairs * 3 * ws.coords.itemsize
    m = int(max(1000, min(ws.npairs, 256e6 / max(per_pair, 1.0))))
    try:
        with tempfile.TemporaryDirectory() as td:
            f = os.path.join(td, "sample.npz")
            np.savez(f, coords=ws.coords, off=ws.off, cnt=ws.cnt, pa=ws.pa[:m], pb=ws.pb[:m])
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_gpu_baseline.py"), "--npz", f] + (["epa"] if epa else []),
                               capture_output=True, text=True, timeout=300)
        out = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:  # a crash of the reference library is a property of the reference, not of this run
        return {"unavailable": f"{type(e).__name__}: {str(e)[:120]}"}
    if "ms" not in out:
        return out
    return {"value": out["pairs"] / out["ms"] * 1e3, "unit": "pairs/s", "sample": f"the first {out['pairs']} pairs of the CPU sample",
            "kind": "reference CUDA kernels, unmodified, kernel-only (device arrays built once), same GPU"}


def config_of(name, n_total, world):
    """The `config` object: identical for our arm and the reference arm."""
    desc = WORKLOADS[name][0]
    return {"workload": desc, "pairs_per_step": int(n_total), "pairs_per_gpu": int(n_total // max(world, 1)),
            "precision": "f64" if WORKLOADS[name][1] == np.float64 else "f32", "seed": 42}


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the SAME workload
    (same generator, seed, pair count per GPU and mix), all host threads, rank 0 only."""
    if rank != 0:
        return
    name = args.workload
    desc, dt, maker, default_pairs = WORKLOADS[name][:4]
    epa = len(WORKLOADS[name]) > 4 and WORKLOADS[name][4]
    prec = "f64" if dt == np.float64 else "f32"
    from oracle.pyoracle import Oracle, Reference
    cores = os.cpu_count() or 1
    n = args.pairs or default_pairs
    # one GPU's share of the job per step (weak scaling: every GPU adds its own n pairs)
    n_step = min(n, args.ref_sample) if args.ref_sample else n
    if maker is None:  # device-generated workloads: the host construction of the same distribution
        maker = (lambda n_, s_: W.random_hulls(n_, 8, 64, s_, 1.0, np.float32, 0.5)) if epa else \
            (lambda n_, s_: W.boxes(n_, True, s_, dtype=np.float32))
        n_step = min(n_step, 2_000_000)
    w = maker(n_step, args.seed)
    if epa:
        kind, o = "port", Oracle(prec)
        run = lambda: o.gjk_epa_batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=cores)
    elif Reference.available(prec):
        kind, ref = "reference", Reference(prec)
        prep = Reference.Prepared(ref, w.coords, w.off, w.cnt, w.pa, w.pb)
        run = lambda: prep.run(cores)
    else:
        kind, o = "port", Oracle(prec)
        run = lambda: o.batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=cores)
    with quiet_stdout():
        for _ in range(args.warmup):
            run()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            run()
        dt_s = time.perf_counter() - t0
    v = n_step * args.steps / dt_s
    sample = (f"the full per-GPU batch ({n_step} pairs) per step" if n_step == n else f"{n_step} of {n} pairs per step")
    line = {
        "impl": "reference", "metric": "gjk_pair_queries_per_sec", "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt_s / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": prec, "data": "synthetic",
        "config": config_of(name, n * max(args.gpus, 1), max(args.gpus, 1)),
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": kind,
                         "sample": f"{sample}, all host threads (OpenMP dynamic split over pairs); one host serves the whole job"},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "unmodified scalar/openGJK.c (oracle/_ref) called once per pair through compute_minimum_distance; "
                "the host's cores do not grow with --gpus, so this rate is the same at every N",
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
    ap.add_argument("--ref-sample", type=int, default=0, help="reference arm: cap on pairs per step (0 = the full batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extra", default=None,
                    help="comma list of further workloads reported under other_workloads "
                         f"(default: {','.join(DEFAULT_EXTRAS)} when --workload is the default; 'none' disables)")
    ap.add_argument("--no-multi", action="store_true", help="skip the one-process MultiEngine measurement at N > 1")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    if args.extra is None:
        extras = list(DEFAULT_EXTRAS) if (args.workload == "hulls64_f64" and not args.pairs) else []
    else:
        extras = [x for x in args.extra.split(",") if x and x != "none"]
    for x in extras:
        if x not in WORKLOADS:
            raise SystemExit(f"unknown workload {x}")

    import torch
    import torch.distributed as dist
    from opengjk_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    cpu_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")  # host-side barrier for the phases where only rank 0 drives the GPUs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    eng = api.Engine(local_rank)
    if os.environ.get("OGJK_CLASSES"):  # e.g. "4:0,24:4,64:8,192:16,2048:32"
        CLASS_TABLE[:] = [tuple(int(x) for x in c.split(":")) for c in os.environ["OGJK_CLASSES"].split(",")]
    eng.set_classes(CLASS_TABLE)
    if os.environ.get("OGJK_ACCEL_MIN"):  # tuning hook: 0 = no cluster tables in the resident store
        eng.set_accel_min(int(os.environ["OGJK_ACCEL_MIN"]))
    if os.environ.get("OGJK_TABLE_LANES"):
        eng.set_table_lanes(int(os.environ["OGJK_TABLE_LANES"]))
    hbm_peak, peak_src = peaks()
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()

    # ---- bare pinned-copy probe: the same H2D + D2H byte counts, no kernels -----------------
    def copy_probe(h2d_bytes, d2h_bytes, reps=3):
        """Wall time of moving h2d_bytes in and d2h_bytes out over this GPU's PCIe link on two
        streams (full duplex), capped at 1 GiB buffers walked repeatedly; every rank at once."""
        cap = 1 << 30
        hb, db = min(int(h2d_bytes), cap), min(int(d2h_bytes), cap)
        h_in = torch.empty(max(hb, 1), dtype=torch.uint8).pin_memory()
        h_out = torch.empty(max(db, 1), dtype=torch.uint8).pin_memory()
        d_in = torch.empty(max(hb, 1), dtype=torch.uint8, device=dev)
        d_out = torch.empty(max(db, 1), dtype=torch.uint8, device=dev)
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

        def once():
            left_in, left_out = int(h2d_bytes), int(d2h_bytes)
            with torch.cuda.stream(s_in):
                while left_in > 0:
                    k = min(left_in, hb)
                    d_in[:k].copy_(h_in[:k], non_blocking=True)
                    left_in -= k
            with torch.cuda.stream(s_out):
                while left_out > 0:
                    k = min(left_out, db)
                    h_out[:k].copy_(d_out[:k], non_blocking=True)
                    left_out -= k
            s_in.synchronize()
            s_out.synchronize()
        once()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            once()
        barrier()
        return max_over_ranks((time.perf_counter() - t0) / reps)

    # ---- device-side generators (inputs never exist on the host) -----------------------------
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

    def device_hulls(n, seed, nmin=8, nmax=64, center_scale=1.0, radius_jitter=0.5):
        """Config #5 inputs generated on the GPU in 1 M-pair slabs (same construction as
        workloads.random_hulls: U{8..64} vertices on a sphere of radius 1 + 0.5u, body 2 centred
        at (u - 0.5) per axis; computed in fp64, rounded once to fp32)."""
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        cnt = torch.randint(nmin, nmax + 1, (2 * n,), generator=g, device=dev, dtype=torch.int32)
        off = torch.cumsum(cnt.to(torch.int64), 0) - cnt
        total = int(off[-1].item()) + int(cnt[-1].item())
        coords = torch.empty(total * 3, dtype=torch.float32, device=dev)
        slab = 1_000_000
        for lo in range(0, n, slab):
            m = min(slab, n - lo)
            k = cnt[2 * lo:2 * (lo + m)].to(torch.int64)
            v0 = int(off[2 * lo].item())
            nv = int(k.sum().item())
            pid = torch.repeat_interleave(torch.arange(2 * m, device=dev), k)
            u = torch.rand((nv, 2), generator=g, device=dev, dtype=torch.float64)
            z = 2.0 * u[:, 0] - 1.0
            th = 2.0 * np.pi * u[:, 1]
            r = torch.sqrt(torch.clamp(1.0 - z * z, min=0.0))
            rad = 1.0 + radius_jitter * torch.rand(2 * m, generator=g, device=dev, dtype=torch.float64)
            centre = torch.zeros((2 * m, 3), device=dev, dtype=torch.float64)
            centre[1::2] = center_scale * (torch.rand((m, 3), generator=g, device=dev, dtype=torch.float64) - 0.5)
            pts = torch.stack([r * torch.cos(th), r * torch.sin(th), z], 1) * rad[pid, None] + centre[pid]
            coords[3 * v0:3 * (v0 + nv)] = pts.to(torch.float32).reshape(-1)
            del pid, u, z, th, r, rad, centre, pts
        pa = torch.arange(0, 2 * n, 2, dtype=torch.int32, device=dev)
        return coords, off, cnt, pa, pa + 1

    def measure(name, pairs, steps, warmup, with_cpu, cpu_budget):
        desc, dt, maker, default_pairs = WORKLOADS[name][:4]
        epa = len(WORKLOADS[name]) > 4 and WORKLOADS[name][4]
        prec = "f64" if dt == np.float64 else "f32"
        sdt = api.SIMPLEX_F64 if prec == "f64" else api.SIMPLEX_F32
        tdt = torch.float64 if prec == "f64" else torch.float32
        strong = name in STRONG
        n_total = (pairs or default_pairs) * (1 if strong else world)
        if strong:
            lo, hi = api.shard_bounds(n_total, world, rank)  # the library's partition of ONE batch
            n = hi - lo
        else:
            n = pairs or default_pairs
        stream = torch.cuda.current_stream().cuda_stream
        w = None
        if maker is not None:
            w = maker(n, args.seed + rank)  # each rank owns its own slice of the job
            h_coords, h_off, h_cnt, h_pa, h_pb = pin(w.coords.reshape(-1)), pin(w.off), pin(w.cnt), pin(w.pa), pin(w.pb)
            d_coords, d_off, d_cnt, d_pa, d_pb = (t.to(dev) for t in (h_coords, h_off, h_cnt, h_pa, h_pb))
            alg_bytes = w.algorithmic_bytes()
        else:
            gen = device_hulls if epa else device_boxes
            d_coords, d_off, d_cnt, d_pa, d_pb = gen(n, args.seed + 1000 * rank)
            verts = d_coords.numel() // 3
            alg_bytes = verts * 3 * 4 + n * (108 + 4)
        in_bytes = int(sum(t.numel() * t.element_size() for t in (d_coords, d_off, d_cnt, d_pa, d_pb)))
        d_dist = torch.empty(n, dtype=tdt, device=dev)
        d_simp = torch.empty(n * sdt.itemsize, dtype=torch.uint8, device=dev)
        if epa:
            d_w1, d_w2, d_nr = (torch.empty(n * 3, dtype=tdt, device=dev) for _ in range(3))
            alg_bytes += n * 36

        # The engine's device-resident format is the packed polytope store (built once, outside the
        # timed region -- the counterpart of the reference's allocate_and_copy_device_arrays); a step
        # is one ogjk_store_gjk_* (ogjk_store_gjk_epa_*) call over it.
        # Built twice: the first build also pays the process's one-off costs (module load, first
        # allocations); the second is what a caller's next pool costs.
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        store = eng.store_create(d_coords, d_off, d_cnt, stream)
        torch.cuda.synchronize()
        store_build_first_ms = 1e3 * (time.perf_counter() - t0)
        store.close()
        t0 = time.perf_counter()
        store = eng.store_create(d_coords, d_off, d_cnt, stream)
        torch.cuda.synchronize()
        store_build_ms = 1e3 * (time.perf_counter() - t0)

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
        ms_max = max_over_ranks(e0.elapsed_time(e1))
        value = n_total * steps / (ms_max * 1e-3)

        # same job starting from the raw xyz-interleaved arrays (repack inside the step)
        raw = None
        if not epa and in_bytes < 8e9:
            for _ in range(2):
                step_from_raw()
            torch.cuda.synchronize()
            reps = max(2, steps // 2)
            e0.record()
            for _ in range(reps):
                step_from_raw()
            e1.record()
            torch.cuda.synchronize()
            raw_ms = e0.elapsed_time(e1) / reps
            raw = {"value": n * world / (raw_ms * 1e-3), "ms_per_step": raw_ms, "note": "repack into the store inside the step"}

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
        # the dominant kernel is charged the whole batch's algorithmic bytes (conservative:
        # other size classes, if any, moved some of them)
        achieved = alg_bytes / (class_avg[dom] * 1e-3) / 1e9
        tname = "double" if prec == "f64" else "float"
        min_cnt = int(d_cnt.min().item())
        if epa and dom == 5:
            kname = f"epa_warp_kernel<{tname}>"
        else:
            lanes = CLASS_TABLE[dom][1]
            amin = int(os.environ.get("OGJK_ACCEL_MIN", 512))
            # as run_store_batch routes: pairs whose two bodies both carry a table run the table kernel,
            # timed with the last class
            if store.table_bytes > 0 and dom == len(CLASS_TABLE) - 1 and min_cnt >= amin:
                # default routing: pairs with a body above 1024 vertices -> tile kernel; the others -> pair-batched
                # kernel in batches of >= 131072 pairs, first-generation kernel below
                tl = int(os.environ.get("OGJK_TABLE_LANES", 0))
                big = int(d_cnt.max().item()) > 1024
                batched = tl == 1 or (not big and (tl == 2 or (tl == 0 and n >= 131072)))  # kBatchKernelMinPairs, abi.cu
                kname = (f"gjk_table3_kernel<{tname}>" if (tl == 3 or (tl in (0, 2) and big))
                         else f"gjk_tablebatch_kernel<{tname}>" if batched
                         else f"gjk_table_kernel<{tname},{tl if tl in (8, 16, 32) else 32}>")
            else:
                kname = f"gjk_small_kernel<{tname}>" if lanes == 0 else f"gjk_group_kernel<{tname},{lanes}>"
        traffic, traffic_src = traffic_for(name)
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": traffic, "traffic_source": traffic_src, "kernel": kname, "kernel_ms": class_avg[dom],
                    "step_phases_ms": {"sort": float(np.mean(sort_ms)), "classes": class_avg},
                    "algorithmic_bytes_per_launch": int(alg_bytes), "peak_source": peak_src}
        if name in NCU_NOTES:
            roofline.update(NCU_NOTES[name])
        else:
            roofline["limiter"] = "HBM-shaped: one compulsory read of both hulls per pair (large hulls)"
        if store.table_bytes > 0:
            # a one-shot caller pays the cluster-table build once per pool
            roofline["one_shot"] = {"store_build_ms": store_build_ms, "first_store_build_in_process_ms": store_build_first_ms,
                                    "pairs_per_s_including_build":
                                    n / ((ms_max / steps + store_build_ms) * 1e-3),
                                    "note": "persistent stores amortise the table build; this is one batch with the build inside"}

        # ---- end to end through the host-buffer C-ABI call --------------------------
        e2e_steps = max(2, min(steps, 5))
        if w is not None:
            m = n
            out_bytes = m * (d_dist.element_size() + sdt.itemsize) + (m * 9 * d_dist.element_size() if epa else 0)
            h_dist = torch.empty(m, dtype=tdt).pin_memory()
            h_simp = torch.empty(m * sdt.itemsize, dtype=torch.uint8).pin_memory()
            o = (h_dist.numpy(), h_simp.numpy().view(sdt))
            hc, ho, hk, ha, hb = (t.numpy() for t in (h_coords, h_off, h_cnt, h_pa, h_pb))
            pair_owned = bool(np.array_equal(w.pa, np.arange(0, 2 * n, 2)) and np.array_equal(w.pb, w.pa + 1))
            e2e_in = in_bytes
            if pair_owned and not epa:
                # the reference's bd1[] / bd2[] call shape: two pinned pools, pair p = polytope p of each
                # (chunk-pipelined H2D / kernels / D2H inside the library)
                pools = [pin(x).numpy() for x in w.two_pools()]
                e2e_call = lambda: eng.gjk_host_pools(*pools, out=o)
                e2e_in = int(sum(x.nbytes for x in pools))
            elif epa:
                e2e_call = lambda: eng.gjk_epa_host(hc, ho, hk, ha, hb)
            else:
                e2e_call = lambda: eng.gjk_host(hc, ho, hk, ha, hb, True, o)
            e2e_note = "the whole batch"
        else:
            # device-generated workloads: a sub-batch is copied to pinned host memory and goes through
            # the host-buffer API (two pools, pair p = polytope p of each)
            m = min(n, 1_000_000 if epa else 8_000_000)
            k_h = d_cnt[:2 * m].cpu().numpy()
            nv = int(k_h.astype(np.int64).sum())
            c_h = d_coords[:3 * nv].cpu().numpy().reshape(-1, 3)
            ws = W.Workload("sub", c_h, W._offsets(k_h), k_h, np.arange(0, 2 * m, 2, dtype=np.int32),
                            np.arange(1, 2 * m, 2, dtype=np.int32))
            pools = [pin(x).numpy() for x in ws.two_pools()]
            e2e_in = int(sum(x.nbytes for x in pools))
            out_bytes = m * (4 + sdt.itemsize) + (m * 36 if epa else 0)
            o = (torch.empty(m, dtype=tdt).pin_memory().numpy(),
                 torch.empty(m * sdt.itemsize, dtype=torch.uint8).pin_memory().numpy().view(sdt))
            if epa:
                eo = tuple(torch.empty((m, 3), dtype=tdt).pin_memory().numpy() for _ in range(3))
                e2e_call = lambda: eng.gjk_host_pools(*pools, out=o, epa=True, epa_out=eo)
            else:
                e2e_call = lambda: eng.gjk_host_pools(*pools, out=o)
            e2e_note = f"a {m}-pair sub-batch copied back from the device-generated inputs"
            del ws, c_h
        e2e_call()  # warm (allocates staging)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_call()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        probe_s = copy_probe(e2e_in, out_bytes)
        e2e_rate = m * world * e2e_steps / e2e_s
        e2e = {"value": e2e_rate, "unit": "pairs/s", "h2d_bytes_per_step": e2e_in, "d2h_bytes_per_step": int(out_bytes),
               "steps": e2e_steps, "pairs_per_step_per_gpu": m, "batch": e2e_note,
               "api": (f"ogjk_gjk_epa_host_{prec}" if epa else f"ogjk_gjk_host_{prec}")
               + " (pinned host buffers in, distances + gkSimplex" + (" + witnesses + normals" if epa else "") + " out)",
               "copy_probe": {"pairs_per_s": m * world / probe_s, "seconds_per_step": probe_s,
                              "h2d_gb_s_per_gpu": e2e_in / probe_s / 1e9,
                              "note": "the same H2D and D2H byte counts on two streams, no kernels, every rank at once"},
               "frac_of_copy_probe": e2e_rate / (m * world / probe_s)}

        res = {"workload": desc, "value": value, "scaling": "strong" if strong else "weak", "from_raw_arrays": raw,
               "store_build_ms": store_build_ms, "store_bytes": store.nbytes, "ms_per_step": ms_max / steps,
               "launches": launches, "roofline": roofline, "e2e": e2e, "clocks": clk.summary(), "prec": prec,
               "pairs_per_gpu": n, "pairs_total": n_total, "input_bytes_per_gpu": in_bytes}

        # ---- parity sample against the oracle + CPU baseline (checker only, outside any timed region) ----
        if rank == 0 and with_cpu:
            from oracle.pyoracle import Oracle
            mchk = min(n, 20000)
            if w is not None:
                wchk = w.slice_pairs(0, mchk)
                cpu_w, cpu_note = w, None
            else:
                k_h = d_cnt[:2 * 100_000].cpu().numpy() if n >= 100_000 else d_cnt.cpu().numpy()
                npoly = k_h.shape[0]
                nv = int(k_h.astype(np.int64).sum())
                c_h = d_coords[:3 * nv].cpu().numpy().reshape(-1, 3)
                cpu_w = W.Workload("device_sample", c_h, W._offsets(k_h), k_h, np.arange(0, npoly, 2, dtype=np.int32),
                                   np.arange(1, npoly, 2, dtype=np.int32))
                wchk = cpu_w.slice_pairs(0, mchk)
                cpu_note = f"the first {npoly // 2} pairs of the device-generated batch (i.i.d. generator)"
            with quiet_stdout():
                if epa:
                    ref = Oracle(prec).gjk_epa_batch(wchk.coords, wchk.off, wchk.cnt, wchk.pa, wchk.pb, nthreads=os.cpu_count() or 1)
                    res["sample_parity"] = {"pairs": mchk, "distance_bits_equal": bool(
                        d_dist[:mchk].cpu().numpy().tobytes() == ref["distances"].tobytes()),
                        "witness_bits_equal": bool(d_w1[:3 * mchk].cpu().numpy().tobytes() == ref["witness1"].tobytes()
                                                   and d_w2[:3 * mchk].cpu().numpy().tobytes() == ref["witness2"].tobytes()),
                        "mean_epa_iterations": float(ref["epa_iters"].mean())}
                else:
                    dref, sref = Oracle(prec).batch(wchk.coords, wchk.off, wchk.cnt, wchk.pa, wchk.pb, nthreads=os.cpu_count() or 1)
                    res["sample_parity"] = {"pairs": mchk, "distance_bits_equal": bool(
                        d_dist[:mchk].cpu().numpy().tobytes() == dref.tobytes()),
                        "simplex_bytes_equal": bool(d_simp[:mchk * sdt.itemsize].cpu().numpy().tobytes() == sref.tobytes())}
            res["cpu_baseline"] = cpu_baseline(cpu_w, prec, cpu_budget, epa, cpu_note)
        store.close()
        del d_coords, d_simp, d_dist, store
        torch.cuda.empty_cache()
        return res

    def multi_engine_run():
        """N > 1: ONE process (rank 0) drives all N GPUs through the product's multi-device entry
        (ogjk_multi_gjk_host_*): the main workload's whole job (N x pairs) in one call, host buffers
        in and out.  The other ranks idle on a host-side barrier."""
        name = args.workload
        desc, dt, maker, default_pairs = WORKLOADS[name][:4]
        if maker is None:
            return None
        out = None
        torch.cuda.synchronize()
        dist.barrier(group=cpu_group)
        if rank == 0:
            try:
                if ORIG_AFFINITY is not None:
                    os.sched_setaffinity(0, ORIG_AFFINITY)  # this process now drives every GPU
                prec = "f64" if dt == np.float64 else "f32"
                sdt = api.SIMPLEX_F64 if prec == "f64" else api.SIMPLEX_F32
                # the whole job of the main workload, capped at 2 M pairs (host generation + three host
                # copies of an 8 x 1 M fp64 job would need > 40 GB and a minute of numpy time)
                n = min((args.pairs or default_pairs) * world, 2_000_000)
                w = maker(n, args.seed)
                pools = [pin(x).numpy() for x in w.two_pools()]
                o = (torch.empty(n, dtype=torch.float64 if prec == "f64" else torch.float32).pin_memory().numpy(),
                     torch.empty(n * sdt.itemsize, dtype=torch.uint8).pin_memory().numpy().view(sdt))
                m = api.MultiEngine(list(range(world)))
                m.gjk_host_pools(*pools, out=o)
                l0 = m.launch_count
                t0 = time.perf_counter()
                reps = 3
                for _ in range(reps):
                    m.gjk_host_pools(*pools, out=o)
                dt_s = (time.perf_counter() - t0) / reps
                launches = (m.launch_count - l0) // reps
                m.close()
                # one-device answer for the same job: byte equality of a strided sample
                idx = np.arange(0, n, max(1, n // 20000))[:20000]
                from oracle.pyoracle import Oracle
                with quiet_stdout():
                    dref, sref = Oracle(prec).batch(w.coords, w.off, w.cnt, w.pa[idx], w.pb[idx], nthreads=os.cpu_count() or 1)
                got_s = o[1][idx]
                bad_d = int((o[0][idx].view(np.uint64 if prec == "f64" else np.uint32) != dref.view(np.uint64 if prec == "f64" else np.uint32)).sum())
                bad_f = {f: int((np.ascontiguousarray(got_s[f]).view(np.uint8).reshape(len(idx), -1)
                                 != np.ascontiguousarray(sref[f]).view(np.uint8).reshape(len(idx), -1)).any(1).sum())
                         for f in ("nvrtx", "vrtx", "vrtx_idx", "witnesses")}
                same = bad_d == 0 and not any(bad_f.values())  # field-wise: struct padding is not part of the contract
                out = {"api": f"ogjk_multi_gjk_host_{prec} (one process, one call, {world} devices; host buffers in and out)",
                       "devices": world, "pairs_per_call": n, "value": n / dt_s, "unit": "pairs/s", "seconds_per_call": dt_s,
                       "h2d_bytes_per_call": int(sum(x.nbytes for x in pools)), "gpu_launches_per_call": int(launches),
                       "sample_parity": {"pairs": int(len(idx)), "bytes_equal_to_oracle": same, "distance_mismatches": bad_d,
                                         "simplex_field_mismatches": bad_f}}
            except Exception as e:  # report, do not lose the main line
                out = {"error": repr(e)}
        dist.barrier(group=cpu_group)
        return out

    with_cpu = not args.no_cpu_baseline
    main_res = measure(args.workload, args.pairs, args.steps, args.warmup, with_cpu, 10.0)
    others = {}
    for name in extras:
        if name == args.workload:
            continue
        r = measure(name, 0, max(3, args.steps // 2), 3, with_cpu, 4.0)
        others[name] = {k: r[k] for k in ("workload", "value", "scaling", "ms_per_step", "launches", "roofline", "e2e", "prec",
                                          "pairs_per_gpu", "pairs_total", "from_raw_arrays", "store_build_ms", "clocks") if k in r}
        for k in ("cpu_baseline", "sample_parity"):
            if k in r:
                others[name][k] = r[k]
    multi = multi_engine_run() if (world > 1 and not args.no_multi) else None

    if rank == 0:
        line = {
            "metric": "gjk_pair_queries_per_sec", "value": main_res["value"], "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": main_res["prec"], "data": "synthetic",
            "config": config_of(args.workload, main_res["pairs_total"], world),
            "run": {"sharding": "contiguous pair slice per rank (ogjk_multi_shard_bounds), no data-path collective",
                    "l2": (f"inputs {main_res['input_bytes_per_gpu'] / 1e6:.0f} MB per GPU "
                           + ("> 126 MB L2 (no flush needed)" if main_res["input_bytes_per_gpu"] > L2_BYTES
                              else "< L2: numbers are L2-warm")),
                    "resident_format": "packed polytope store (ogjk_store_create, outside the timed region)",
                    "host_binding": numa or "unbound",
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
        if multi is not None:
            line["multi_engine"] = multi
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
