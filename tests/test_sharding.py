"""Multi-GPU host logic on CPU: two gloo ranks shard a pair list with the library's own
partition function (contiguous slice per rank, no data-path collective), each computes its slice
(with the CPU oracle standing in for the device), and an all_gather reassembles
results identical to the single-process run.  Also checks the max-over-ranks timing
reduction bench.py uses."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard_bounds(npairs, world, rank):
    """Contiguous slice of the pair list owned by `rank`: the LIBRARY's partition
    (ogjk_multi_shard_bounds, csrc/multi.cuh) -- the same function the multi-GPU entry points
    and bench.py use.  Pure host arithmetic, callable without a GPU."""
    sys.path.insert(0, ROOT)
    from opengjk_b200 import api
    return api.shard_bounds(npairs, world, rank)


def _worker(rank, world, port, npairs, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from opengjk_b200 import workloads as W
    from oracle.pyoracle import Oracle
    w = W.mixed_hulls(npairs, seed=3)
    lo, hi = shard_bounds(npairs, world, rank)
    d, s = Oracle("f64").batch(w.coords, w.off, w.cnt, w.pa[lo:hi], w.pb[lo:hi])
    # optional final gather of results (the only communication; sizes differ per rank)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([hi - lo]))
    cap = int(max(x.item() for x in sizes))
    pad = torch.zeros(cap, dtype=torch.float64)
    pad[: hi - lo] = torch.from_numpy(d)
    parts = [torch.zeros(cap, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(parts, pad)
    # bench.py's timing reduction: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        full = np.concatenate([parts[r][: int(sizes[r].item())].numpy() for r in range(world)])
        np.save(out_path, np.concatenate([full, [t.item()]]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_exactly():
    for n in (0, 1, 7, 1000, 1001, 100_000_000, 2**31 + 5):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_two_rank_gloo_sharding_matches_single_process(tmp_path):
    npairs, world = 2001, 2
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(world, port, npairs, out), nprocs=world, join=True)
    got = np.load(out)
    sys.path.insert(0, ROOT)
    from opengjk_b200 import workloads as W
    from oracle.pyoracle import Oracle
    w = W.mixed_hulls(npairs, seed=3)
    d, _ = Oracle("f64").batch(w.coords, w.off, w.cnt, w.pa, w.pb)
    assert got[:-1].tobytes() == d.tobytes()
    assert got[-1] == float(world)  # max over ranks
