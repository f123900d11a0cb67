"""EPA pinned to the reference: tests/golden/epa_ref_cuda_{f32,f64}.npz hold outputs of the
reference's own CUDA GJK+EPA (gpu/opengjk_gpu.cu compiled unmodified on a B200 by
tests/golden/make_epa_golden.py, once with its default flags and once with -fmad=false).  The
reference has no CPU EPA and no numeric EPA test, so these files are the oracle's pin.

Bar: against the no-FMA build the CPU restatement (and, on a GPU, our CUDA kernels) reproduce
distance and both witness points BIT FOR BIT; against the default (FMA-contracting) build the
penetration depth agrees within the north-star tolerance (1e-5 fp32 / 1e-10 fp64) on every pair
(witnesses of the FMA build differ on tie-sensitive pairs -- the two reference builds disagree
with each other there in the same way)."""
import os

import numpy as np
import pytest

from oracle.pyoracle import Oracle, simplex_dtype
from tests.golden.make_epa_golden import checksum, epa_workloads

GOLD = os.path.join(os.path.dirname(__file__), "golden")
PRECS = [("f64", np.float64, 1e-10, np.uint64), ("f32", np.float32, 1e-5, np.uint32)]
# EPA outputs (penetrating pairs) are bit-equal on EVERY pair.  For SEPARATED pairs EPA only copies the
# GJK result, and the golden files hold the reference's CUDA GJK there, whose support reduction
# (gpu/opengjk_gpu.cu:1229-1241: __shfl_down tree over lane chunks, "other > mine" strict) resolves an exact
# tie between two vertices of non-adjacent chunks towards whatever the tree leaves in lane 0 -- not the lowest
# index the scalar reference takes (scalar/openGJK.c:617-622).  Root-caused on f64 hulls_mixed pair 306
# (iteration 5: vertices 3 and 14 of body 1 have the identical dot 0x1.e04ed492d452ap-3; scalar -> 3,
# CUDA -> 14; traced with instrumented copies of both reference sources) and pair 1442.  This repo's GJK
# follows the scalar reference (the north-star contract), so those pairs differ from the CUDA golden in the
# last bits; at most 0.5 % of the separated pairs (2 of 512 here) may do so, all within the tolerance.
MIN_BIT_EQUAL_SEPARATED = 0.995


def _bits(a, it):
    return np.ascontiguousarray(a).view(it)


def _compare(name, prec, tol, it, g, dist, w1, w2):
    for variant in ("", "_nofma"):
        k = name + variant
        if k + "/failed" in g.files:
            # the reference kernel itself faults on this input (fp32 axis-aligned cubes: CUDA error 700)
            assert name == "cubes_aligned_overlap" and prec == "f32"
            continue
        assert bool(g[k + "/repeatable"])
        d, r1, r2 = g[k + "/dist"], g[k + "/w1"], g[k + "/w2"]
        assert np.array_equal(d < 0, dist < 0), k                                  # same pairs penetrate
        assert np.isclose(d, dist, rtol=tol, atol=tol).all(), k                    # depth / distance: every pair
        if variant == "_nofma":
            deq = _bits(d, it) == _bits(dist, it)
            weq = (_bits(r1, it) == _bits(w1, it)).all(1) & (_bits(r2, it) == _bits(w2, it)).all(1)
            pen = d < 0
            assert deq[pen].all() and weq[pen].all(), (k, "EPA outputs", deq[pen].mean(), weq[pen].mean())
            if (~pen).any():
                assert deq[~pen].mean() >= MIN_BIT_EQUAL_SEPARATED and weq[~pen].mean() >= MIN_BIT_EQUAL_SEPARATED, \
                    (k, "separated pairs (reference CUDA GJK tie rule)", deq[~pen].mean(), weq[~pen].mean())
            assert np.isclose(r1, w1, rtol=tol, atol=tol).all() and np.isclose(r2, w2, rtol=tol, atol=tol).all(), k


@pytest.mark.parametrize("prec,dt,tol,it", PRECS)
def test_cpu_restatement_matches_reference_cuda_outputs(prec, dt, tol, it):
    g = np.load(os.path.join(GOLD, f"epa_ref_cuda_{prec}.npz"))
    o = Oracle(prec)
    for name, w in epa_workloads(dt):
        assert checksum(w) == bytes(g[name + "/sha"]).decode()
        r = o.gjk_epa_batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=os.cpu_count() or 1)
        _compare(name, prec, tol, it, g, r["distances"], r["witness1"], r["witness2"])


@pytest.mark.parametrize("prec,dt,tol,it", PRECS)
def test_epa_stage_alone_on_reference_gjk_results(prec, dt, tol, it):
    """The reference's compute_epa fed with the reference's own GPU GJK simplices: the EPA
    restatement reproduces depth, witnesses and contact normal of every penetrating pair exactly."""
    g = np.load(os.path.join(GOLD, f"epa_ref_cuda_{prec}.npz"))
    w = dict(epa_workloads(dt))["boxes_overlap"]
    k = "boxes_overlap_nofma"
    gs = g[k + "/gjk_simplex"].view(simplex_dtype(prec)).reshape(-1)
    gd = g[k + "/gjk_dist"]
    r = Oracle(prec).gjk_epa_batch(w.coords, w.off, w.cnt, w.pa, w.pb, simplices=gs, distances=gd)
    # pairs the EPA gate (distance <= eps, gpu/opengjk_gpu.cu:2053) lets through; the reference's stand-alone
    # compute_epa does not upload the caller's distances, so its output for separated pairs is not meaningful
    pen = ~(gd > np.finfo(dt).eps)
    assert pen.mean() > 0.9 and (g[k + "/epa_dist"][pen] < 0).all()
    for ref, got in ((g[k + "/epa_dist"], r["distances"]), (g[k + "/epa_w1"], r["witness1"]),
                     (g[k + "/epa_w2"], r["witness2"])):
        assert (_bits(ref, it)[pen] == _bits(got, it)[pen]).all()
    assert np.array_equal(g[k + "/epa_normals"][pen], r["normals"][pen])            # (== ignores the sign of zero)


@pytest.mark.gpu
@pytest.mark.parametrize("prec,dt,tol,it", PRECS)
def test_cuda_kernels_match_reference_cuda_outputs(prec, dt, tol, it):
    from opengjk_b200 import api
    eng = api.Engine(0)
    g = np.load(os.path.join(GOLD, f"epa_ref_cuda_{prec}.npz"))
    for name, w in epa_workloads(dt):
        r = eng.gjk_epa_host(w.coords, w.off, w.cnt, w.pa, w.pb)
        _compare(name, prec, tol, it, g, r["distances"], r["witness1"], r["witness2"])
    eng.close()


def test_tie_resolution_follows_the_reference_shuffle_tree():
    """An exact tie between lanes is resolved by the reference's __shfl_down tree
    (gpu/opengjk_gpu.cu:1780-1801), not by the lowest index: with 8 vertices (one per lane) whose
    vertices 1 and 4 coincide at the extreme point, lane 0 takes lane 4's candidate at offset 4 and
    keeps it when it meets lane 1 at offset 1 (strict >), so the answer is 4 where a first-index
    argmax says 1.  Overlapping boxes hit such ties constantly; the restatement emulates the tree."""
    import ctypes as C
    o = Oracle("f64")
    fn = o.lib.oracle_epa_support_f64
    pts = np.array([[0.1, 0.2, 0.0], [2.0, 0.0, 0.0], [0.3, -0.4, 0.1], [0.5, 0.5, 0.5],
                    [2.0, 0.0, 0.0], [-0.2, 0.1, 0.3], [0.0, -0.6, 0.2], [-1.0, 0.0, 0.0]], dtype=np.float64)
    other = np.array([[0.0, 0.0, 0.0], [0.5, 0.1, 0.0], [-3.0, 0.2, 0.1], [0.2, 0.9, -0.3]], dtype=np.float64)
    out = np.zeros(3); idx = np.zeros(2, dtype=np.int32)
    d = np.array([1.0, 0.0, 0.0])
    fn(C.c_int(8), pts.ctypes.data_as(C.c_void_p), C.c_int(4), other.ctypes.data_as(C.c_void_p),
       d.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p))
    assert int(np.argmax(pts @ d)) == 1          # the textbook rule
    assert idx[0] == 4 and idx[1] == 2, idx      # the reference's rule; body 2 along -d has a unique maximum
    assert np.array_equal(out, pts[4] - other[2])
