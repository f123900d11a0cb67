"""Reference-named batched flavour (libopenGJK_GPU.so / _f64.so): the host symbols of
gpu/examples/python/openGJK_GPU.def with the reference's signatures, called the way
pyopengjk_gpu calls them (arrays of gkPolytope with flattened host `coord`), checked
against the CPU oracle.  Covers the high-level, indexed, EPA and mid-level APIs."""
import ctypes as C
import os

import numpy as np
import pytest

from opengjk_b200 import workloads as W
from oracle.pyoracle import Oracle, polytope_dtype, simplex_dtype

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBS = {"f32": os.path.join(ROOT, "opengjk_b200", "libopenGJK_GPU.so"),
        "f64": os.path.join(ROOT, "opengjk_b200", "libopenGJK_GPU_f64.so")}
DEF_SYMBOLS = ["compute_minimum_distance", "compute_epa", "compute_gjk_epa", "compute_minimum_distance_indexed",
               "compute_minimum_distance_device", "compute_epa_device", "allocate_and_copy_device_arrays",
               "allocate_epa_device_arrays", "copy_results_from_device", "copy_epa_results_from_device",
               "free_device_arrays", "free_epa_device_arrays"]


def test_flavour_library_exports_the_def_list():
    for path in LIBS.values():
        lib = C.CDLL(path)
        for sym in DEF_SYMBOLS:
            assert hasattr(lib, sym), (path, sym)


def _bodies(w, which, prec):
    arr = np.zeros(which.shape[0], dtype=polytope_dtype(prec))
    arr["numpoints"] = w.cnt[which]
    arr["coord"] = w.coords.ctypes.data + w.off[which].astype(np.uint64) * np.uint64(3 * w.coords.itemsize)
    return arr


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.gpu
@pytest.mark.parametrize("prec,dt", [("f32", np.float32), ("f64", np.float64)])
def test_high_level_and_indexed_calls(prec, dt):
    lib = C.CDLL(LIBS[prec])
    w = W.mixed_hulls(5000, seed=71, dtype=dt)
    dref, sref = Oracle(prec).batch(w.coords, w.off, w.cnt, w.pa, w.pb)
    n = w.npairs
    b1, b2 = _bodies(w, w.pa, prec), _bodies(w, w.pb, prec)
    simp = np.zeros(n, dtype=simplex_dtype(prec))
    dist = np.zeros(n, dtype=dt)
    lib.compute_minimum_distance(C.c_int(n), _p(b1), _p(b2), _p(simp), _p(dist))
    assert dist.tobytes() == dref.tobytes()
    for f in ("nvrtx", "vrtx", "vrtx_idx", "witnesses"):
        assert np.ascontiguousarray(simp[f]).tobytes() == np.ascontiguousarray(sref[f]).tobytes(), f
    # indexed: one pool + gkCollisionPair[]
    pool = _bodies(w, np.arange(w.npoly), prec)
    pairs = np.ascontiguousarray(np.stack([w.pa, w.pb], 1).astype(np.int32))
    dist2 = np.zeros(n, dtype=dt)
    simp2 = np.zeros(n, dtype=simplex_dtype(prec))
    lib.compute_minimum_distance_indexed(C.c_int(w.npoly), C.c_int(n), _p(pool), _p(pairs), _p(simp2), _p(dist2))
    assert dist2.tobytes() == dref.tobytes()
    # n <= 0 is a silent no-op (gpu/opengjk_gpu.cu:2984)
    lib.compute_minimum_distance(C.c_int(0), None, None, None, None)


@pytest.mark.gpu
@pytest.mark.parametrize("prec,dt", [("f32", np.float32), ("f64", np.float64)])
def test_gjk_epa_and_mid_level_api(prec, dt):
    lib = C.CDLL(LIBS[prec])
    w = W.random_hulls(3000, 8, 64, 72, 1.5, dt, 0.5)
    ref = Oracle(prec).gjk_epa_batch(w.coords, w.off, w.cnt, w.pa, w.pb)
    n = w.npairs
    b1, b2 = _bodies(w, w.pa, prec), _bodies(w, w.pb, prec)
    sdt = simplex_dtype(prec)
    simp, dist = np.zeros(n, dtype=sdt), np.zeros(n, dtype=dt)
    w1, w2 = np.zeros((n, 3), dtype=dt), np.zeros((n, 3), dtype=dt)
    lib.compute_gjk_epa(C.c_int(n), _p(b1), _p(b2), _p(simp), _p(dist), _p(w1), _p(w2))
    assert dist.tobytes() == ref["distances"].tobytes()
    assert w1.tobytes() == ref["witness1"].tobytes() and w2.tobytes() == ref["witness2"].tobytes()

    # mid level: allocate -> GJK on device -> EPA on device -> copy back -> free (openGJK_GPU.h:286-399)
    vp = C.c_void_p
    d = [vp() for _ in range(6)]
    lib.allocate_and_copy_device_arrays(C.c_int(n), _p(b1), _p(b2), *[C.byref(x) for x in d])
    e = [vp() for _ in range(3)]
    lib.allocate_epa_device_arrays(C.c_int(n), *[C.byref(x) for x in e])
    lib.compute_minimum_distance_device(C.c_int(n), d[0], d[1], d[4], d[5])
    simp2, dist2 = np.zeros(n, dtype=sdt), np.zeros(n, dtype=dt)
    lib.copy_results_from_device(C.c_int(n), d[4], d[5], _p(simp2), _p(dist2))
    gd, gs = Oracle(prec).batch(w.coords, w.off, w.cnt, w.pa, w.pb)
    assert dist2.tobytes() == gd.tobytes()
    assert np.ascontiguousarray(simp2["vrtx_idx"]).tobytes() == np.ascontiguousarray(gs["vrtx_idx"]).tobytes()
    lib.compute_epa_device(C.c_int(n), d[0], d[1], d[4], d[5], e[0], e[1], e[2])
    w1b, w2b, nrb = (np.zeros((n, 3), dtype=dt) for _ in range(3))
    lib.copy_results_from_device(C.c_int(n), d[4], d[5], _p(simp2), _p(dist2))
    lib.copy_epa_results_from_device(C.c_int(n), e[0], e[1], e[2], _p(w1b), _p(w2b), _p(nrb))
    assert dist2.tobytes() == ref["distances"].tobytes()
    assert w1b.tobytes() == ref["witness1"].tobytes() and nrb.tobytes() == ref["normals"].tobytes()
    lib.free_device_arrays(*d)
    lib.free_epa_device_arrays(*e)


@pytest.mark.gpu
def test_eight_concurrent_callers_on_the_batched_flavour():
    """The flavour library's single context is created once (std::call_once) and calls are
    serialised: 8 threads calling compute_minimum_distance at the same time (the first call of
    the process included) all get the oracle's bytes."""
    import subprocess
    import sys
    code = r'''
import ctypes as C, os, sys, threading
import numpy as np
sys.path.insert(0, %r)
from opengjk_b200 import workloads as W
from oracle.pyoracle import Oracle, polytope_dtype, simplex_dtype
lib = C.CDLL(%r)
prec, dt = "f32", np.float32
ws = [W.mixed_hulls(3000, seed=200 + t, dtype=dt) for t in range(8)]
refs = [Oracle(prec).batch(w.coords, w.off, w.cnt, w.pa, w.pb) for w in ws]
out = [None] * 8
def body(w, which):
    arr = np.zeros(which.shape[0], dtype=polytope_dtype(prec))
    arr["numpoints"] = w.cnt[which]
    arr["coord"] = w.coords.ctypes.data + w.off[which].astype(np.uint64) * np.uint64(3 * w.coords.itemsize)
    return arr
start = threading.Barrier(8)
def work(t):
    w = ws[t]
    b1, b2 = body(w, w.pa), body(w, w.pb)
    simp = np.zeros(w.npairs, dtype=simplex_dtype(prec)); dist = np.zeros(w.npairs, dtype=dt)
    start.wait()
    for _ in range(5):
        lib.compute_minimum_distance(C.c_int(w.npairs), b1.ctypes.data_as(C.c_void_p), b2.ctypes.data_as(C.c_void_p),
                                     simp.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p))
    out[t] = (dist, simp)
th = [threading.Thread(target=work, args=(t,)) for t in range(8)]
[x.start() for x in th]; [x.join() for x in th]
for t in range(8):
    assert out[t][0].tobytes() == refs[t][0].tobytes(), t
    assert np.ascontiguousarray(out[t][1]["vrtx_idx"]).tobytes() == np.ascontiguousarray(refs[t][1]["vrtx_idx"]).tobytes(), t
print("CONCURRENT_OK")
''' % (ROOT, LIBS["f32"])
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert "CONCURRENT_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


@pytest.mark.gpu
def test_devices_environment_switch_spreads_the_drop_in_call():
    """OPENGJK_B200_DEVICES=all: the flavour library's compute_minimum_distance / compute_gjk_epa run on every
    GPU of the box (ogjk_multi_*) and still return the oracle's bytes; a bad value falls back to one device."""
    import subprocess
    import sys
    code = r'''
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, %r)
from opengjk_b200 import workloads as W
from oracle.pyoracle import Oracle, polytope_dtype, simplex_dtype
lib = C.CDLL(%r)
prec, dt = "f32", np.float32
p = lambda a: a.ctypes.data_as(C.c_void_p)
def body(w, which):
    arr = np.zeros(which.shape[0], dtype=polytope_dtype(prec))
    arr["numpoints"] = w.cnt[which]
    arr["coord"] = w.coords.ctypes.data + w.off[which].astype(np.uint64) * np.uint64(3 * w.coords.itemsize)
    return arr
w = W.mixed_hulls(30001, seed=301, dtype=dt)
dref, sref = Oracle(prec).batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=os.cpu_count() or 1)
b1, b2 = body(w, w.pa), body(w, w.pb)
simp = np.zeros(w.npairs, dtype=simplex_dtype(prec)); dist = np.zeros(w.npairs, dtype=dt)
lib.compute_minimum_distance(C.c_int(w.npairs), p(b1), p(b2), p(simp), p(dist))
assert dist.tobytes() == dref.tobytes()
assert np.ascontiguousarray(simp["vrtx_idx"]).tobytes() == np.ascontiguousarray(sref["vrtx_idx"]).tobytes()
we = W.random_hulls(4001, 8, 64, 302, 1.0, dt, 0.5)
eref = Oracle(prec).gjk_epa_batch(we.coords, we.off, we.cnt, we.pa, we.pb, nthreads=os.cpu_count() or 1)
b1, b2 = body(we, we.pa), body(we, we.pb)
simp = np.zeros(we.npairs, dtype=simplex_dtype(prec)); dist = np.zeros(we.npairs, dtype=dt)
w1 = np.zeros((we.npairs, 3), dtype=dt); w2 = np.zeros((we.npairs, 3), dtype=dt)
lib.compute_gjk_epa(C.c_int(we.npairs), p(b1), p(b2), p(simp), p(dist), p(w1), p(w2))
assert dist.tobytes() == eref["distances"].tobytes()
assert w1.tobytes() == np.ascontiguousarray(eref["witness1"]).tobytes()
print("DEVICES_OK")
''' % (ROOT, LIBS["f32"])
    for value in ("all", "0", "not-a-device-list"):
        env = dict(os.environ, OPENGJK_B200_DEVICES=value)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
        assert "DEVICES_OK" in r.stdout, value + ": " + r.stdout[-2000:] + r.stderr[-3000:]
