"""Reference-named scalar flavour (libopengjk_ce.so): same symbols, by-value struct
signature and ctypes binding pattern as the reference's pyopengjk
(scalar/examples/python_ctypes/src/pyopengjk/opengjk.py:7-91), so its test-suite
style carries over.  Also drives the device sub-solver through the reference's
opengjk_test_S{1,2,3}D hooks against the reference's known-answer vectors."""
import ctypes
import json
import os
from ctypes import POINTER, Structure, c_double, c_float, c_int

import numpy as np
import pytest

from opengjk_b200 import api, workloads as W
from oracle.pyoracle import Oracle
from tests.test_oracle import ANALYTIC, _close

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB64 = os.path.join(ROOT, "opengjk_b200", "libopengjk_ce.so")
LIB32 = os.path.join(ROOT, "opengjk_b200", "libopengjk_ce_f32.so")
KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "subsolver_kat.json")))


def binding(path, cf):
    class POLYTOPE(Structure):  # opengjk.py:7-13
        _fields_ = [("numpoints", c_int), ("s", cf * 3), ("s_idx", c_int), ("coord", POINTER(POINTER(cf)))]

    class SIMPLEX(Structure):  # opengjk.py:16-22
        _fields_ = [("nvrtx", c_int), ("vrtx", (cf * 3) * 4), ("vrtx_idx", (c_int * 2) * 4),
                    ("witnesses", (cf * 3) * 2)]

    lib = ctypes.cdll.LoadLibrary(path)
    lib.compute_minimum_distance.restype = cf
    lib.compute_minimum_distance.argtypes = [POLYTOPE, POLYTOPE, POINTER(SIMPLEX)]

    def query(v0, v1):  # opengjk.py:60-91
        bodies = []
        for verts in (v0, v1):
            p = POLYTOPE()
            p.numpoints = len(verts)
            p.s = (cf * 3)(0, 0, 0)
            p.coord = (POINTER(cf) * len(verts))()
            for i, vertex in enumerate(verts):
                p.coord[i] = (cf * 3)(*[float(x) for x in vertex])
            bodies.append(p)
        s = SIMPLEX()
        s.nvrtx = 0
        d = lib.compute_minimum_distance(bodies[0], bodies[1], s)
        return d, s

    return lib, query


def test_flavour_libraries_export_the_reference_symbols():
    for path in (LIB64, LIB32):
        lib = ctypes.CDLL(path)
        for sym in ("compute_minimum_distance", "opengjk_test_S1D", "opengjk_test_S2D", "opengjk_test_S3D", "csFunction"):
            assert hasattr(lib, sym), (path, sym)
    hdr = open(os.path.join(ROOT, "include", "openGJK", "openGJK.h")).read()
    for sym in ("compute_minimum_distance", "opengjk_test_S1D", "opengjk_test_S2D", "opengjk_test_S3D", "gkPolytope",
                "gkSimplex", "USE_32BITS"):
        assert sym in hdr


@pytest.mark.gpu
def test_example_two_bodies_by_value_call():
    _, q64 = binding(LIB64, c_double)
    d, s = q64(W.USER_P, W.USER_Q)
    assert f"{d:f}" == "3.653650" and d == 3.6536497222945008  # README.md:55-59, SURVEY 8(c) golden
    assert s.nvrtx == 2 and [tuple(s.vrtx_idx[i]) for i in range(2)] == [(6, 6), (1, 1)]
    _, q32 = binding(LIB32, c_float)
    d32, _ = q32(W.USER_P, W.USER_Q)
    assert np.float32(d32) == np.float32(3.6536495685577393)


@pytest.mark.gpu
def test_csharp_entry_point():
    """csFunction (scalar/openGJK.c:1154-1220, the P/Invoke target of scalar/examples/cs/main.cs):
    component-major coordinate arrays in, distance out."""
    for path, cf, want in ((LIB64, c_double, 3.6536497222945008), (LIB32, c_float, float(np.float32(3.6536495685577393)))):
        lib = ctypes.CDLL(path)
        lib.csFunction.restype = cf
        lib.csFunction.argtypes = [c_int, POINTER(cf), c_int, POINTER(cf)]
        a = np.ascontiguousarray(np.asarray(W.USER_P, dtype=cf).T)   # (3, n): component-major
        b = np.ascontiguousarray(np.asarray(W.USER_Q, dtype=cf).T)
        d = lib.csFunction(a.shape[1], a.ctypes.data_as(POINTER(cf)), b.shape[1], b.ctypes.data_as(POINTER(cf)))
        assert d == want


@pytest.mark.gpu
@pytest.mark.parametrize("case", ANALYTIC[::3], ids=lambda c: c[0])
def test_reference_analytic_cases_through_the_drop_in(case):
    _, q64 = binding(LIB64, c_double)
    _, a, b, exp, wscale = case
    d, s = q64(a, b)
    assert np.isclose(d, exp, atol=1e-12)
    if wscale is not None:
        w0 = np.array(s.witnesses[0][:]); w1 = np.array(s.witnesses[1][:])
        assert np.linalg.norm(w0 - w1) == pytest.approx(d, abs=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("case", KAT["sub"], ids=lambda c: c["src"])
def test_device_subsolver_known_answers(case):
    v, s = api.subalgorithm(case["vrtx"], dtype=np.float64)
    if case["nvrtx"] is not None:
        assert int(s["nvrtx"]) == case["nvrtx"]
    for got, want in zip(v, case["v"]):
        if want is not None:
            assert _close(float(got), want), (v, case["v"])


@pytest.mark.gpu
def test_device_subsolver_matches_oracle_bitwise():
    rng = np.random.default_rng(5)
    for prec, dt in (("f64", np.float64), ("f32", np.float32)):
        o = Oracle(prec)
        for n in (2, 3, 4):
            for _ in range(300):
                pts = rng.normal(size=(n, 3)) * rng.choice([1e-3, 1.0, 50.0])
                if rng.random() < 0.3:
                    pts[:, rng.integers(3)] *= 1e-9
                idx = rng.integers(0, 50, size=(n, 2))
                vo, so = o.sub(pts, idx)
                vd, sd = api.subalgorithm(pts, idx, dt)
                assert vo.tobytes() == vd.tobytes()
                for f in ("nvrtx", "vrtx", "vrtx_idx"):
                    assert so[f].tobytes() == sd[f].tobytes(), f


@pytest.mark.gpu
def test_eight_concurrent_callers_on_the_scalar_flavour():
    """The reference's compute_minimum_distance is re-entrant; the drop-in serialises callers on
    a mutex around its lazily created context.  8 threads (racing on the first call of the
    process) must all get the reference's answers."""
    import subprocess
    import sys
    code = r'''
import sys, threading
import numpy as np
sys.path.insert(0, %r)
from ctypes import c_double
from tests.test_compat_scalar import binding, LIB64
from opengjk_b200 import workloads as W
from oracle.pyoracle import Oracle
_, q = binding(LIB64, c_double)
w = W.mixed_hulls(160, seed=300, dtype=np.float64)
dref, sref = Oracle("f64").batch(w.coords, w.off, w.cnt, w.pa, w.pb)
got = np.zeros(w.npairs)
idx = np.zeros((w.npairs, 4, 2), dtype=np.int32)
start = threading.Barrier(8)
def work(t):
    start.wait()
    for p in range(t, w.npairs, 8):
        d, s = q(w.polytope(w.pa[p]), w.polytope(w.pb[p]))
        got[p] = d
        idx[p] = np.array([[s.vrtx_idx[i][0], s.vrtx_idx[i][1]] for i in range(4)])
th = [threading.Thread(target=work, args=(t,)) for t in range(8)]
[x.start() for x in th]; [x.join() for x in th]
assert got.tobytes() == dref.tobytes()
assert idx.tobytes() == np.ascontiguousarray(sref["vrtx_idx"]).tobytes()
print("CONCURRENT_OK")
''' % (ROOT,)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert "CONCURRENT_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
