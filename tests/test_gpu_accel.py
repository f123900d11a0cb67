"""GPU parity of the cluster-table support scan (persistent stores, large polytopes -- the
tests lower the threshold to 256 vertices): pruning must not change a single output byte.
Every case is compared with the CPU oracle AND with the same store built without
tables."""
import os

import numpy as np
import pytest

from opengjk_b200 import api, workloads as W
from oracle.pyoracle import Oracle
from tests.test_gpu_parity import _assert_same

pytestmark = pytest.mark.gpu
PRECS = [("f64", np.float64), ("f32", np.float32)]


@pytest.fixture(scope="module")
def eng():
    e = api.Engine(0)
    yield e
    e.set_accel_min(512)
    e.set_table_lanes(0)
    e.close()


def _store_run(eng, w, accel_min, lanes=0):
    import torch
    prec = "f64" if w.coords.dtype == np.float64 else "f32"
    sdt = api.SIMPLEX_F64 if prec == "f64" else api.SIMPLEX_F32
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    eng.set_accel_min(accel_min)
    eng.set_table_lanes(lanes)
    store = eng.store_create(w.coords, w.off, w.cnt)
    nbytes = store.nbytes
    pa, pb = t(w.pa), t(w.pb)
    simp = torch.zeros(w.npairs * sdt.itemsize, dtype=torch.uint8, device=dev)
    dist = torch.zeros(w.npairs, dtype=torch.float64 if prec == "f64" else torch.float32, device=dev)
    eng.store_gjk(store, pa, store, pb, w.npairs, simp, dist, 0)
    torch.cuda.synchronize()
    store.close()
    return dist.cpu().numpy(), simp.cpu().numpy().view(sdt).reshape(-1), nbytes


def _check(eng, w, prec, expect_tables=True):
    o = Oracle(prec)
    dref, sref = o.batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=os.cpu_count() or 1)
    d0, s0, plain_bytes = _store_run(eng, w, 0)
    _assert_same(w.name + "/plain", d0, s0, dref, sref, prec)
    # 0: default routing; 1: pair-batched kernel (32 pairs per warp) for every table pair; 2: tile kernel above
    # 1024 vertices + pair-batched kernel below; 3: tile kernel for every table pair; 8/16/32: first-generation
    # kernel with that many lanes per pair
    for lanes in (0, 1, 2, 3, 8, 16, 32):
        d1, s1, table_bytes = _store_run(eng, w, 256, lanes)
        _assert_same(f"{w.name}/tables/lanes{lanes}", d1, s1, dref, sref, prec)
        assert (table_bytes > plain_bytes) == expect_tables, (plain_bytes, table_bytes)


def _workload(name, bodies, pairs, dt):
    cnt = np.array([len(b) for b in bodies], dtype=np.int32)
    off = np.concatenate([[0], np.cumsum(cnt)[:-1]]).astype(np.int64)
    pts = np.concatenate(bodies).astype(dt)
    pa = np.array([p[0] for p in pairs], dtype=np.int32)
    pb = np.array([p[1] for p in pairs], dtype=np.int32)
    return W.Workload(name, pts, off, cnt, pa, pb)


@pytest.mark.parametrize("prec,dt", PRECS)
def test_sphere_hulls_1k(eng, prec, dt):
    _check(eng, W.large_hull_pool(4000, 128, 1000, 1000, 5, dtype=dt), prec)


@pytest.mark.parametrize("prec,dt", PRECS)
def test_ragged_sizes_around_threshold(eng, prec, dt):
    # 200..700 vertices: some polytopes below the threshold, counts not multiples of 32 or 4
    _check(eng, W.large_hull_pool(4000, 200, 200, 700, 6, dtype=dt), prec)


@pytest.mark.parametrize("prec,dt", PRECS)
def test_large_and_oversized(eng, prec, dt):
    # up to 10k vertices (many table tiles per scan) and two polytopes above the 16384 cap
    w = W.large_hull_pool(300, 24, 3000, 10000, 7, dtype=dt)
    _check(eng, w, prec)
    rng = np.random.default_rng(8)
    big = [rng.normal(size=(n, 3)) for n in (20000, 17000, 5000, 300)]
    big = [b / np.linalg.norm(b, axis=1, keepdims=True) + c for b, c in zip(big, ([0, 0, 0], [3, 0, 0], [0, 2.5, 0], [1, 1, 1]))]
    _check(eng, _workload("oversized", big, [(0, 1), (0, 2), (1, 2), (2, 3), (0, 3), (3, 1)], dt), prec)


@pytest.mark.parametrize("prec,dt", PRECS)
def test_exact_ties_lattice_bodies(eng, prec, dt):
    """Integer lattice points on cube surfaces: almost every support query has many exactly
    equal dot products, so the lowest-ORIGINAL-index rule is what decides."""
    rng = np.random.default_rng(11)
    g = np.arange(9) - 4.0
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    shell = np.stack([x, y, z], -1).reshape(-1, 3)
    shell = shell[np.abs(shell).max(1) == 4]          # 386 surface points
    bodies, pairs = [], []
    for k in range(40):
        perm = rng.permutation(len(shell))
        rep = np.concatenate([shell[perm], shell[rng.integers(0, len(shell), 50)]])  # + exact duplicates
        shift = np.round(rng.uniform(-12, 12, 3)) if k % 2 else np.zeros(3)
        bodies.append(rep + shift)
    for k in range(2000):
        a, b = rng.integers(0, 40, 2)
        pairs.append((a, b))
    _check(eng, _workload("lattice", bodies, pairs, dt), prec)


@pytest.mark.parametrize("prec,dt", PRECS)
def test_degenerate_large_bodies(eng, prec, dt):
    rng = np.random.default_rng(12)
    same = np.tile([[1.0, 2.0, 3.0]], (300, 1))                       # all vertices identical
    line = np.outer(np.linspace(-1, 1, 513), [1.0, 0.5, 0.25])        # collinear
    flat = np.c_[rng.uniform(-1, 1, (400, 2)), np.zeros(400)]         # coplanar
    blob = rng.normal(size=(1000, 3))                                 # interior points too
    sph = rng.normal(size=(777, 3)); sph /= np.linalg.norm(sph, axis=1, keepdims=True)
    small = rng.normal(size=(10, 3)) + [4, 0, 0]
    bodies = [same, line, flat, blob, sph, small, sph + [0.5, 0.5, 0], blob * 1e-3, sph * 1e3 + [5e3, 0, 0]]
    pairs = [(i, j) for i in range(len(bodies)) for j in range(len(bodies))]
    _check(eng, _workload("degenerate", bodies, pairs * 8, dt), prec)


def test_non_finite_coordinates(eng):
    """NaN / inf vertices: outputs still equal the oracle's position-wise (NaN == NaN)."""
    rng = np.random.default_rng(13)
    a = rng.normal(size=(500, 3)); a /= np.linalg.norm(a, axis=1, keepdims=True)
    b = a + [3.0, 0, 0]
    a_nan = a.copy(); a_nan[137, 1] = np.nan
    b_inf = b.copy(); b_inf[400, 0] = np.inf
    w = _workload("nonfinite", [a, b, a_nan, b_inf], [(0, 1), (2, 1), (0, 3), (2, 3), (1, 2), (3, 0)] * 16, np.float64)
    o = Oracle("f64")
    dref, sref = o.batch(w.coords, w.off, w.cnt, w.pa, w.pb)
    d1, s1, _ = _store_run(eng, w, 256)
    assert d1.tobytes() == dref.tobytes()
    for f in ("nvrtx", "vrtx", "vrtx_idx", "witnesses"):
        assert np.ascontiguousarray(s1[f]).tobytes() == np.ascontiguousarray(sref[f]).tobytes(), f


def test_tables_with_epa(eng):
    """The GJK+EPA store entry on a store that carries tables (EPA reads the plain blocks)."""
    import torch
    dt = np.float32
    w = W.large_hull_pool(512, 32, 300, 600, 14, radius=1.0, offset_scale=1.0, dtype=dt)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    outs = []
    for amin in (0, 256):
        eng.set_accel_min(amin)
        store = eng.store_create(w.coords, w.off, w.cnt)
        n = w.npairs
        simp = torch.zeros(n * api.SIMPLEX_F32.itemsize, dtype=torch.uint8, device=dev)
        dist = torch.zeros(n, dtype=torch.float32, device=dev)
        w1, w2, nr = (torch.zeros(n * 3, dtype=torch.float32, device=dev) for _ in range(3))
        eng.store_gjk_epa(store, t(w.pa), store, t(w.pb), n, simp, dist, w1, w2, nr, 0)
        torch.cuda.synchronize()
        outs.append([x.cpu().numpy().tobytes() for x in (simp, dist, w1, w2, nr)])
        store.close()
    assert outs[0] == outs[1]


@pytest.mark.parametrize("prec,dt", PRECS)
def test_two_stores_unindexed_and_mixed_thresholds(eng, prec, dt):
    """Un-indexed form over two different stores (pair p = polytope p of each), the stores built
    with different table thresholds; one store without tables at all falls back to plain scans."""
    import torch
    w = W.large_hull_pool(1500, 64, 300, 1500, 15, dtype=dt)
    dref, sref = Oracle(prec).batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=os.cpu_count() or 1)
    c1, o1, k1, c2, o2, k2 = w.two_pools()
    sdt = api.SIMPLEX_F64 if prec == "f64" else api.SIMPLEX_F32
    dev = torch.device("cuda:0")
    for amin1, amin2 in ((256, 512), (512, 0), (0, 0), (64, 1024)):
        eng.set_accel_min(amin1)
        s1 = eng.store_create(c1, o1, k1)
        eng.set_accel_min(amin2)
        s2 = eng.store_create(c2, o2, k2)
        assert (s1.table_bytes > 0) == (amin1 > 0) and (s2.table_bytes > 0) == (amin2 > 0)
        simp = torch.zeros(w.npairs * sdt.itemsize, dtype=torch.uint8, device=dev)
        dist = torch.zeros(w.npairs, dtype=torch.float64 if prec == "f64" else torch.float32, device=dev)
        eng.store_gjk(s1, None, s2, None, w.npairs, simp, dist, 0)
        torch.cuda.synchronize()
        _assert_same(f"two_stores/{amin1}/{amin2}", dist.cpu().numpy(), simp.cpu().numpy().view(sdt).reshape(-1), dref, sref, prec)
        s1.close(); s2.close()
