"""Pins the CPU oracle (oracle/gjk_oracle.c) to the reference.

Runs without a GPU.  Sources of truth, in the order SURVEY.md 8(c) lists them:
  * the reference's sub-solver known-answer vectors (tests/golden/subsolver_kat.json,
    extracted from scalar/tests/*.c),
  * the reference's analytic end-to-end cases, re-stated from
    scalar/examples/python_ctypes/test/test_min_distance.py (same geometry, same
    1e-12 tolerance),
  * byte-equality with outputs of the unmodified reference: committed goldens
    (tests/golden/gjk_golden_*.npz) and, when oracle/_ref is present, live runs.
"""
import json
import math
import os

import numpy as np
import pytest

from opengjk_b200 import workloads as W
from oracle.pyoracle import Oracle, Reference
from tests.golden.make_golden import checksum, golden_workloads

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-12  # test_min_distance.py:8-9


@pytest.fixture(scope="module")
def o64():
    return Oracle("f64")


@pytest.fixture(scope="module")
def o32():
    return Oracle("f32")


KAT = json.load(open(os.path.join(GOLD, "subsolver_kat.json")))


def _close(a, b):
    # cmocka's assert_float_equal: absolute eps or relative FLT_EPSILON
    return abs(a - b) <= max(1e-12, 1.2e-7 * max(abs(a), abs(b)))


@pytest.mark.parametrize("case", KAT["hff"], ids=lambda c: c["src"])
def test_hff_known_answers(o64, case):
    assert o64.hff(case["which"], *case["args"]) == case["expect"]


@pytest.mark.parametrize("case", KAT["sub"], ids=lambda c: c["src"])
def test_subsolver_known_answers(o64, case):
    v, s = o64.sub(case["vrtx"])
    if case["nvrtx"] is not None:
        assert int(s["nvrtx"]) == case["nvrtx"]
    for got, want in zip(v, case["v"]):
        if want is not None:
            assert _close(float(got), want), (v, case["v"])


def test_subsolver_matches_reference_bitwise():
    if not Reference.available("f64"):
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(3)
    for prec in ("f64", "f32"):
        o, r = Oracle(prec), Reference(prec)
        for n in (2, 3, 4):
            for _ in range(3000):
                pts = rng.normal(size=(n, 3)) * rng.choice([1e-3, 1.0, 50.0])
                if rng.random() < 0.3:  # near-degenerate: flatten one axis
                    pts[:, rng.integers(3)] *= 1e-9
                idx = rng.integers(0, 50, size=(n, 2))
                vo, so = o.sub(pts, idx)
                vr, sr = r.sub(pts, idx)
                assert vo.tobytes() == vr.tobytes()
                assert so.tobytes() == sr.tobytes()


# ---- analytic cases (reference test_min_distance.py) -------------------------
def _pt_line(p1, p2, q):
    return np.linalg.norm(np.cross(p2 - p1, p1 - q)) / np.linalg.norm(p2 - p1)


def _pt_plane(p1, p2, p3, q):
    n = np.cross(p2 - p1, p3 - p1)
    return abs(np.dot(n / np.linalg.norm(n), q - p2))


def _check_witness(d, s, scale=1.0):
    gap = np.linalg.norm(np.array(s["witnesses"][0], np.float64) - np.array(s["witnesses"][1], np.float64))
    assert gap == pytest.approx(d, abs=max(1e-5, scale * 1e-6))


def analytic_cases():
    """(name, bodyA, bodyB, expected distance or None, witness scale or None)"""
    out = []
    line = np.array([[0.1, 0.2, 0.3], [0.5, 0.8, 0.7]])
    for delta in (0.1, 1e-12, 0, -2):  # :46-56
        p = line[0] + 0.27 * (line[1] - line[0]) + delta * np.cross(line[0], line[1])
        out.append((f"line_point[{delta}]", line, p[None], _pt_line(line[0], line[1], p), 1.0))
    l1 = np.array([[-0.5, -0.7, -0.3], [1, 2, 3]], dtype=np.float64)
    for delta in (0.1, 1e-12, 0):  # :59-71
        p = l1[0] + 0.38 * (l1[1] - l1[0]) + delta * np.cross(l1[0], l1[1])
        l2 = np.array([p, [2, 5, 6]])
        out.append((f"line_line[{delta}]", l1, l2, _pt_line(l1[0], l1[1], l2[0]), 1.0))
    for i in range(6):  # :74-86
        delta = 0.1 ** (3 * i)
        t1 = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float64)
        t2 = np.array([[1, delta, 0], [3, 1.2, 0], [1, 1, 0]], dtype=np.float64)
        out.append((f"tri[{i}]", t1, t2, _pt_line(t1[2], t1[1], t2[0]), 1.0))
    for i in range(6):  # :89-103
        delta = 0.1 * 0.1 ** (3 * i)
        q1 = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], dtype=np.float64)
        q2 = np.array([[0, 1 + delta, 0], [2, 2, 0], [2, 4, 0], [4, 4, 0]], dtype=np.float64)
        out.append((f"quad[{i}]", q1, q2, _pt_line(q1[2], q1[3], q2[0]), 1.0))
    for i in range(7):  # :106-118
        delta = 0.5 ** (3 * i)
        a = np.array([[0, 0, 0.2], [1, 0, 0.1], [0, 1, 0.3], [0, 0, 1]], dtype=np.float64)
        b = np.array([[0, 0, -3], [1, 0, -3], [0, 1, -3], [0.5, 0.3, -delta]], dtype=np.float64)
        out.append((f"tetra[{i}]", a, b, _pt_plane(a[0], a[1], a[2], b[3]), 1.0))
    for i in range(6):  # :121-137
        delta = (-1) ** i * np.sqrt(2) * 0.1 ** (3 * i)
        a = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float64)
        b = np.array([[0, 0, -3], [1, 0, -3], [0, 1, -3], [0.5, 0.3, -delta]], dtype=np.float64)
        exp = 0.0 if delta < 0 else _pt_plane(a[0], a[1], a[2], b[3])
        out.append((f"tetra_collision[{i}]", a, b, exp, None))
    h1 = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [0, 0, 1], [1, 0, 1], [0, 1, 1], [1, 1, 1]],
                  dtype=np.float64)
    for delta in (0, -0.1, -0.49, -0.51):  # :140-168
        p0 = np.array([1.5 + delta, 1.5 + delta, 0.5])
        p1 = np.array([2, 2, 1.0])
        p2 = np.array([2, 1.25, 0.25])
        quad = np.array([p0, p1, p2, p1 + p2 - p0])
        n = np.cross(quad[1] - quad[0], quad[2] - quad[0])
        n /= np.linalg.norm(n)
        h2 = np.concatenate([quad, quad + n])
        exp = 0.0 if p0[0] < 1 else np.linalg.norm(np.array([1, 1, p0[2]]) - h2[0])
        out.append((f"hex[{delta}]", h1, h2, exp, 1.0))
    # Go binding KAT (scalar/examples/go/openGJK/connector_test.go:11-33 style: planar quads, exact 0/1/2)
    sq = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], dtype=np.float64)
    for shift, exp in ((0.5, 0.0), (2.0, 1.0), (3.0, 2.0)):
        out.append((f"quads_shift[{shift}]", sq, sq + np.array([shift, 0, 0]), exp, None))
    return out


ANALYTIC = analytic_cases()


@pytest.mark.parametrize("case", ANALYTIC, ids=lambda c: c[0])
def test_analytic_distances(o64, case):
    _, a, b, exp, wscale = case
    d, s, _ = o64.gjk(a, b)
    assert np.isclose(d, exp, atol=TOL)
    if wscale is not None:
        _check_witness(d, s, wscale)


def rotated_cubes():
    from scipy.spatial.transform import Rotation as R
    cubes = [np.array([[-1, -1, -1], [1, -1, -1], [-1, 1, -1], [1, 1, -1], [-1, -1, 1], [1, -1, 1], [-1, 1, 1],
                       [1, 1, 1]], dtype=np.float64)]
    cubes.append(R.from_euler("z", 45, degrees=True).apply(cubes[0]))
    cubes.append(R.from_euler("y", np.arctan2(1.0, np.sqrt(2))).apply(cubes[1]))
    cubes.append(R.from_euler("y", 45, degrees=True).apply(cubes[0]))
    return cubes


@pytest.mark.parametrize("c0", range(4))
@pytest.mark.parametrize("c1", range(4))
def test_cube_distance(o64, c0, c1):  # test_min_distance.py:171-194
    cubes = rotated_cubes()
    dx = cubes[c0][:, 0].max() - cubes[c1][:, 0].min()
    for delta in (1e8, 1.0, 1e-4, 1e-8, 1e-12):
        d, s, _ = o64.gjk(cubes[c0], cubes[c1] + np.array([dx + delta, 0, 0]))
        _check_witness(d, s, max(1.0, delta))
        assert np.isclose(d, delta)


@pytest.mark.parametrize("scale", [1.0, 1e4])
def test_random_tiny_bodies_witness_consistency(o64, scale):  # :197-223
    w = W.tiny_bodies(60, 99, scale)
    d, s = o64.batch(w.coords, w.off, w.cnt, w.pa, w.pb)
    ok = ~np.isnan(d)
    gap = np.linalg.norm(s["witnesses"][:, 0].astype(np.float64) - s["witnesses"][:, 1], axis=1)
    assert np.all(np.abs(gap[ok] - d[ok]) <= max(1e-5, scale * 1e-6))


def test_config1_goldens(o64, o32):
    """SURVEY 8(c) 17-digit goldens for userP/userQ (README.md:55-59 prints 3.653650)."""
    d, s, _ = o64.gjk(W.USER_P, W.USER_Q)
    assert d == 3.6536497222945008 and f"{d:f}" == "3.653650"
    assert int(s["nvrtx"]) == 2 and s["vrtx_idx"][:2].tolist() == [[6, 6], [1, 1]]
    np.testing.assert_array_equal(s["witnesses"][0], [1.0251728907330566, 1.4903181189488242, 0.2554633471645919])
    np.testing.assert_array_equal(s["witnesses"][1], [-1.0251728907330566, -1.4903181189488242, -0.2554633471645919])
    d32, s32, _ = o32.gjk(W.USER_P, W.USER_Q)
    assert np.float32(d32) == np.float32(3.6536495685577393)
    assert s32["vrtx_idx"][:2].tolist() == [[6, 6], [1, 1]]


@pytest.mark.parametrize("prec,dt", [("f64", np.float64), ("f32", np.float32)])
def test_oracle_matches_committed_reference_outputs(prec, dt):
    gold = np.load(os.path.join(GOLD, f"gjk_golden_{prec}.npz"))
    o = Oracle(prec)
    for w in golden_workloads(dt):
        assert checksum(w) == bytes(gold[w.name + "/sha"]).decode(), f"inputs of {w.name} drifted"
        d, s = o.batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=4)
        assert d.tobytes() == gold[w.name + "/dist"].tobytes(), w.name
        assert s.view(np.uint8).reshape(len(s), -1).tobytes() == gold[w.name + "/simplex"].tobytes(), w.name


@pytest.mark.parametrize("prec,dt", [("f64", np.float64), ("f32", np.float32)])
def test_oracle_matches_live_reference(prec, dt):
    if not Reference.available(prec):
        pytest.skip("oracle/_ref not built (reference tree not mounted)")
    o, r = Oracle(prec), Reference(prec)
    for w in (W.mixed_hulls(20000, seed=5, dtype=dt), W.boxes(20000, False, 6, dtype=dt),
              W.boxes(20000, True, 7, dtype=dt), W.tiny_bodies(100, 8, 1.0, dt)):
        do, so = o.batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=4)
        dr, sr = r.batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=4)
        assert do.tobytes() == dr.tobytes()
        assert so.tobytes() == sr.tobytes()


def test_threaded_batch_equals_serial(o64):
    w = W.mixed_hulls(5000, seed=11)
    d1, s1 = o64.batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=1)
    d4, s4 = o64.batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=4)
    assert d1.tobytes() == d4.tobytes() and s1.tobytes() == s4.tobytes()


def test_iteration_cap_never_hit_on_synthetic(o64):
    w = W.mixed_hulls(20000, seed=2)
    _, _, it = o64.batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=4, want_iters=True)
    assert it.max() < 25 and it.min() >= 1


def _cap_workload(dt, n=3000, seed=11):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cap_pair.npz"))
    return W.cap_stress(n, seed, dt, base=(g["a"], g["b"]))


@pytest.mark.parametrize("prec,dt", [("f64", np.float64), ("f32", np.float32)])
def test_iteration_cap_exit_restatement_equals_reference(prec, dt):
    """The k == mk exit (scalar/openGJK.c:1036-1042: witnesses from a non-converged simplex):
    >= 1000 pairs that run all 25 iterations, restatement byte-equal to the unmodified reference."""
    w = _cap_workload(dt)
    o = Oracle(prec)
    d, s, it = o.batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=os.cpu_count() or 1, want_iters=True)
    assert int((it >= 25).sum()) >= 1000, int((it >= 25).sum())
    if not Reference.available(prec):
        pytest.skip("oracle/_ref not built")
    import contextlib
    with open(os.devnull, "w") as devnull:
        saved = os.dup(1)
        os.dup2(devnull.fileno(), 1)   # the reference prints a banner per cap hit
        try:
            dr, sr = Reference(prec).batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=os.cpu_count() or 1)
        finally:
            os.dup2(saved, 1)
            os.close(saved)
    assert d.tobytes() == dr.tobytes()
    for f in ("nvrtx", "vrtx", "vrtx_idx", "witnesses"):
        assert np.ascontiguousarray(s[f]).tobytes() == np.ascontiguousarray(sr[f]).tobytes(), f
