"""GPU EPA: the warp-per-pair shared-memory kernel against the sequential CPU
restatement (oracle/epa_oracle.c), bit for bit, plus analytic depths and the
reference suite's loose range checks (gpu/examples/python/test_examples.py)."""
import os

import numpy as np
import pytest

from opengjk_b200 import api, workloads as W
from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu
PRECS = [("f64", np.float64), ("f32", np.float32)]


@pytest.fixture(scope="module")
def eng():
    e = api.Engine(0)
    yield e
    e.close()


def _same(name, got, ref):
    for k in ("distances", "witness1", "witness2", "normals"):
        a, b = np.ascontiguousarray(got[k]), np.ascontiguousarray(ref[k])
        bad = np.nonzero((a.view(np.uint8).reshape(len(a), -1) != b.view(np.uint8).reshape(len(b), -1)).any(1))[0]
        assert bad.size == 0, f"{name}/{k}: {bad.size} pairs differ, first {bad[:5]}: {a[bad[:2]]} vs {b[bad[:2]]}"
    assert np.array_equal(got["simplices"]["nvrtx"], ref["simplices"]["nvrtx"])


@pytest.mark.parametrize("prec,dt", PRECS)
@pytest.mark.parametrize("maker", [
    lambda dt: W.random_hulls(20000, 8, 64, 7, 1.0, dt, 0.5),       # config #5: all intersecting
    lambda dt: W.boxes(20000, True, 3, 2.0, dt),                     # overlapping oriented boxes
    lambda dt: W.boxes(20000, False, 4, 2.0, dt),                    # axis-aligned (face-face contacts, ties)
    lambda dt: W.mixed_hulls(20000, seed=5, dtype=dt),               # separated + intersecting mix
    lambda dt: W.tiny_bodies(100, 6, 1.0, dt),                       # degenerate 1..7 point bodies
    lambda dt: W.large_hull_pool(500, 32, 200, 1500, 8, 2.0, 1.5, dt),  # large overlapping hulls
], ids=["hulls_intersecting", "boxes_oriented", "boxes_aligned", "mixed", "tiny", "large"])
def test_gjk_epa_matches_oracle_bitwise(eng, prec, dt, maker):
    w = maker(dt)
    ref = Oracle(prec).gjk_epa_batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=os.cpu_count() or 1)
    got = eng.gjk_epa_host(w.coords, w.off, w.cnt, w.pa, w.pb)
    _same(w.name, got, ref)


@pytest.mark.parametrize("prec,dt", PRECS)
def test_store_level_epa_and_epa_only(eng, prec, dt):
    import torch
    w = W.random_hulls(5000, 8, 64, 17, 1.5, dt, 0.5)
    ref = Oracle(prec).gjk_epa_batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=os.cpu_count() or 1)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    sdt = api.SIMPLEX_F64 if prec == "f64" else api.SIMPLEX_F32
    store = eng.store_create(w.coords, w.off, w.cnt)
    pa, pb = t(w.pa), t(w.pb)
    n = w.npairs
    simp = torch.zeros(n * sdt.itemsize, dtype=torch.uint8, device=dev)
    tt = t(w.coords[:1]).dtype
    dist = torch.zeros(n, dtype=tt, device=dev)
    w1, w2, nr = (torch.zeros(n * 3, dtype=tt, device=dev) for _ in range(3))
    st = torch.cuda.current_stream().cuda_stream
    eng.store_gjk_epa(store, pa, store, pb, n, simp, dist, w1, w2, nr, st)
    torch.cuda.synchronize()
    got = {"distances": dist.cpu().numpy(), "witness1": w1.cpu().numpy().reshape(n, 3),
           "witness2": w2.cpu().numpy().reshape(n, 3), "normals": nr.cpu().numpy().reshape(n, 3),
           "simplices": simp.cpu().numpy().view(sdt).reshape(-1)}
    _same("store_gjk_epa", got, ref)
    # EPA only, from GJK results (ogjk_compute_epa_* struct-array form)
    gd, gs = eng.run_workload(w)
    A = [w.polytope(h) for h in w.pa]
    B = [w.polytope(h) for h in w.pb]
    d2, a1, a2, n2 = api.compute_epa_structs(A, B, gs, gd, dt, eng)
    _same("compute_epa", {"distances": d2, "witness1": a1, "witness2": a2, "normals": n2, "simplices": gs}, ref)
    d3, s3, b1, b2 = api.compute_gjk_epa_structs(A, B, dt, eng)
    assert d3.tobytes() == ref["distances"].tobytes() and b1.tobytes() == ref["witness1"].tobytes()
    store.close()


def _cube(side=2.0):
    h = side / 2
    return np.array([[sx * h, sy * h, sz * h] for sz in (-1, 1) for sy in (-1, 1) for sx in (-1, 1)], dtype=np.float64)


@pytest.mark.parametrize("prec,dt,tol", [("f64", np.float64, 1e-9), ("f32", np.float32, 1e-4)])
def test_analytic_box_penetration(eng, prec, dt, tol):
    """Axis-aligned cubes of side 2 overlapping along x by `depth`: EPA depth, normal +-x,
    witnesses |w1-w2| = depth."""
    depths = np.array([0.1, 0.5, 0.9, 1.3])
    A = np.stack([_cube()] * len(depths))
    B = np.stack([_cube() + np.array([2.0 - d, 0.15, -0.1]) for d in depths])
    r = api.compute_epa(A, B, True, dt, eng)
    assert np.allclose(-r["penetration_depths"], depths, atol=tol)
    assert np.allclose(np.abs(r["contact_normals"][:, 0]), 1.0, atol=tol)
    # face-face contact: the witness pair is not unique sideways, but its separation along the normal is the depth
    assert np.allclose(np.abs(r["witnesses1"][:, 0] - r["witnesses2"][:, 0]), depths, atol=10 * tol)


def test_reference_suite_ranges(eng):
    """Loose range checks of gpu/examples/python/test_examples.py (fp32, seed 42)."""
    rng = np.random.default_rng(42)
    cube = _cube().astype(np.float32)
    # touching cubes :267-306: 0 <= d < 0.01
    r = api.compute_minimum_distance(cube[None], (cube + np.float32([2.0, 0, 0]))[None], engine=eng)
    assert 0 <= r["distances"][0] < 0.01
    # overlapping cubes :309-349: depth in (-1.2, -0.8)
    r = api.compute_gjk_epa(cube[None], (cube + np.float32([1.0, 0, 0]))[None], engine=eng)
    assert -1.2 < r["distances"][0] < -0.8
    # separated cubes :352-393: d in (2.9, 3.1), nvrtx < 4
    r = api.compute_gjk_epa(cube[None], (cube + np.float32([5.0, 0, 0]))[None], engine=eng)
    assert 2.9 < r["distances"][0] < 3.1 and r["simplex_nvrtx"][0] < 4
    # overlapping 1000-point spheres :396-463: witnesses within radius + 0.1
    def sphere(n, radius, centre):
        p = rng.normal(size=(n, 3))
        return (radius * p / np.linalg.norm(p, axis=1, keepdims=True) + centre).astype(np.float32)
    s1, s2 = sphere(1000, 1.0, [0, 0, 0]), sphere(1000, 1.0, [1.0, 0, 0])
    r = api.compute_gjk_epa(s1[None], s2[None], engine=eng)
    assert r["simplex_nvrtx"][0] >= 3 and r["distances"][0] < 0
    assert np.linalg.norm(r["witnesses1"][0]) <= 1.1 and np.linalg.norm(r["witnesses2"][0] - [1.0, 0, 0]) <= 1.1
    assert -1.1 < r["distances"][0] < -0.9


def test_loose_agreement_with_reference_cuda_library(eng):
    """The reference's own CUDA kernels (compiled unmodified into oracle/_ref, FMA
    contraction on) run live on the same GPU: GJK distances agree to fp32 rounding and EPA
    depths for the overwhelming majority of pairs.  (The exact pin is tests/test_epa_golden.py:
    committed outputs of the reference's -fmad=false build, matched bit for bit.)"""
    from oracle.pyoracle import ReferenceGPU
    if not ReferenceGPU.available("f32"):
        pytest.skip("oracle/_ref/libopenGJK_GPU_ref_f32.so not shipped")
    ref = ReferenceGPU("f32")
    w = W.mixed_hulls(20000, seed=21, dtype=np.float32)
    dr, sr = ref.gjk(w.coords, w.off, w.cnt, w.pa, w.pb)
    d, s = eng.run_workload(w)
    close = np.isclose(d, dr, rtol=1e-4, atol=1e-5)
    assert close.mean() > 0.999, close.mean()
    assert (s["nvrtx"] == sr["nvrtx"]).mean() > 0.97
    w = W.random_hulls(5000, 8, 64, 23, 1.0, np.float32, 0.5)
    dr, sr, w1r, w2r = ref.gjk(w.coords, w.off, w.cnt, w.pa, w.pb, epa=True)
    got = eng.gjk_epa_host(w.coords, w.off, w.cnt, w.pa, w.pb)
    ok = np.isclose(got["distances"], dr, rtol=1e-3, atol=1e-4)
    assert ok.mean() > 0.97, ok.mean()
