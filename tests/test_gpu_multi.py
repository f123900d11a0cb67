"""Multi-GPU product entry points (csrc/multi.cuh, api.MultiEngine): one call, one batch,
contiguous pair slice per device, results in disjoint ranges of the caller's arrays.
Every result must equal, byte for byte, what one device (and the oracle) gives.
With a single visible GPU the same code runs with ndev = 1 and with the same device
listed through explicit ordinals; the >= 2-device cases skip."""
import ctypes as C
import os

import numpy as np
import pytest

from opengjk_b200 import api, workloads as W
from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu
PRECS = [("f64", np.float64), ("f32", np.float32)]


def _ndev():
    import torch
    return torch.cuda.device_count()


def _cudart():
    """The CUDA runtime torch has loaded (to read back raw device pointers in the resident test)."""
    import torch
    torch.cuda.init()
    for line in open("/proc/self/maps"):
        if "libcudart" in line:
            return C.CDLL(line.split()[-1])
    raise RuntimeError("libcudart not mapped")


def _devsets():
    n = _ndev()
    sets = [[0]]
    if n >= 2:
        sets += [[0, 1], [1, 0]]
    if n >= 4:
        sets += [[0, 1, 2, 3]]
    if n >= 3:
        sets += [list(range(n))]
    return sets


def _same(name, d, s, dref, sref):
    assert d.tobytes() == dref.tobytes(), f"{name}: distance bits differ"
    for f in ("nvrtx", "vrtx", "vrtx_idx", "witnesses"):
        assert np.ascontiguousarray(s[f]).tobytes() == np.ascontiguousarray(sref[f]).tobytes(), (name, f)


@pytest.mark.parametrize("prec,dt", PRECS)
def test_multi_host_entry_points_match_the_oracle(prec, dt):
    o = Oracle(prec)
    w = W.mixed_hulls(50001, seed=91, dtype=dt)           # odd size: slices differ by one pair
    wl = W.large_hull_pool(3001, 48, 300, 1500, 92, dtype=dt)  # indexed, pool replicated
    dref, sref = o.batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=os.cpu_count() or 1)
    dlref, slref = o.batch(wl.coords, wl.off, wl.cnt, wl.pa, wl.pb, nthreads=os.cpu_count() or 1)
    for devs in _devsets():
        m = api.MultiEngine(devs)
        try:
            assert m.ndev == len(devs)
            d, s = m.gjk_host_pools(*w.two_pools())          # un-indexed: sliced pools
            _same(f"pools{devs}", d, s, dref, sref)
            d, s = m.gjk_host(w.coords, w.off, w.cnt, w.pa, w.pb)  # indexed over the pair-owned pool
            _same(f"indexed{devs}", d, s, dref, sref)
            d, s = m.gjk_host(wl.coords, wl.off, wl.cnt, wl.pa, wl.pb)
            _same(f"pool{devs}", d, s, dlref, slref)
            assert m.launch_count > 0
        finally:
            m.close()


@pytest.mark.parametrize("prec,dt", PRECS)
def test_multi_host_chunk_pipelined_slices(prec, dt):
    """Slices large enough for every device to take the chunk-pipelined host path (>= 8 x 4096 pairs per
    device) concurrently from its own host thread, pinned and pageable outputs."""
    import torch
    o = Oracle(prec)
    for devs in _devsets():
        n = 40000 * len(devs) + 7
        w = W.mixed_hulls(n, seed=97, dtype=dt)
        dref, sref = o.batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=os.cpu_count() or 1)
        pools = w.two_pools()
        sdt = api.SIMPLEX_F64 if prec == "f64" else api.SIMPLEX_F32
        m = api.MultiEngine(devs)
        try:
            for pinned in (False, True):
                if pinned:
                    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
                    args = [pin(x) for x in pools]
                    out = (torch.empty(n, dtype=torch.float64 if prec == "f64" else torch.float32).pin_memory().numpy(),
                           torch.empty(n * sdt.itemsize, dtype=torch.uint8).pin_memory().numpy().view(sdt))
                else:
                    args, out = list(pools), None
                for _ in range(2):
                    d, s = m.gjk_host_pools(*args, out=out)
                    _same(f"chunked{devs}/pinned{pinned}", d, s, dref, sref)
        finally:
            m.close()


@pytest.mark.parametrize("prec,dt", PRECS)
def test_multi_epa_and_replicated_stores(prec, dt):
    o = Oracle(prec)
    w = W.random_hulls(20001, 8, 64, 93, 1.0, dt, 0.5)
    ref = o.gjk_epa_batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=os.cpu_count() or 1)
    wl = W.large_hull_pool(4001, 64, 400, 1400, 94, dtype=dt)
    dlref, slref = o.batch(wl.coords, wl.off, wl.cnt, wl.pa, wl.pb, nthreads=os.cpu_count() or 1)
    for devs in _devsets():
        m = api.MultiEngine(devs)
        try:
            got = m.gjk_host_pools(*w.two_pools(), epa=True)
            for k in ("distances", "witness1", "witness2", "normals"):
                assert np.ascontiguousarray(got[k]).tobytes() == np.ascontiguousarray(ref[k]).tobytes(), (devs, k)
            for via_nccl in ([False, True] if len(devs) > 1 else [False]):
                st = m.store_create(wl.coords, wl.off, wl.cnt, via_nccl=via_nccl)
                assert st.nbytes >= len(devs) * wl.coords.nbytes
                d, s = m.store_gjk(st, wl.pa, st, wl.pb)
                _same(f"store{devs}/nccl{via_nccl}", d, s, dlref, slref)
                st.close()
            sw = m.store_create(w.coords, w.off, w.cnt)
            got = m.store_gjk(sw, w.pa, sw, w.pb, epa=True)
            for k in ("distances", "witness1", "witness2", "normals"):
                assert np.ascontiguousarray(got[k]).tobytes() == np.ascontiguousarray(ref[k]).tobytes(), (devs, "store", k)
            sw.close()
        finally:
            m.close()


@pytest.mark.parametrize("prec,dt", PRECS)
def test_multi_struct_array_entries(prec, dt):
    """The reference's call shape (arrays of gkPolytope pointing at host vertex arrays) over several devices:
    ogjk_multi_compute_minimum_distance_* / ogjk_multi_compute_gjk_epa_* equal the oracle byte for byte,
    with a batch large enough for the pipelined gather and one smaller than the device count."""
    o = Oracle(prec)
    w = W.mixed_hulls(40003, seed=97, dtype=dt)
    dref, sref = o.batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=os.cpu_count() or 1)
    we = W.random_hulls(5001, 8, 64, 98, 1.0, dt, 0.5)
    eref = o.gjk_epa_batch(we.coords, we.off, we.cnt, we.pa, we.pb, nthreads=os.cpu_count() or 1)
    bodies = lambda w, which: [w.coords[w.off[h]:w.off[h] + w.cnt[h]] for h in which]
    for devs in _devsets():
        m = api.MultiEngine(devs)
        try:
            d, s = m.compute_minimum_distance_structs(bodies(w, w.pa), bodies(w, w.pb), dt)
            _same(f"structs{devs}", d, s, dref, sref)
            d, s = m.compute_minimum_distance_structs(bodies(w, w.pa[:1]), bodies(w, w.pb[:1]), dt)
            _same(f"structs-one{devs}", d, s, dref[:1], sref[:1])
            got = m.compute_minimum_distance_structs(bodies(we, we.pa), bodies(we, we.pb), dt, epa=True)
            for k in ("distances", "witness1", "witness2"):
                assert np.ascontiguousarray(got[k]).tobytes() == np.ascontiguousarray(eref[k]).tobytes(), (devs, k)
        finally:
            m.close()


def test_resident_results_and_device_side_gather():
    """Results left on the devices: each device's slice equals the oracle's; after the NCCL
    gather every device holds the whole batch (skips the gather with a single GPU)."""
    import torch
    dt, prec = np.float32, "f32"
    wl = W.large_hull_pool(5003, 64, 300, 900, 95, dtype=dt)
    dref, sref = Oracle(prec).batch(wl.coords, wl.off, wl.cnt, wl.pa, wl.pb, nthreads=os.cpu_count() or 1)
    n = wl.npairs
    for devs in _devsets():
        m = api.MultiEngine(devs)
        try:
            st = m.store_create(wl.coords, wl.off, wl.cnt)
            for gather in ([False, True] if len(devs) > 1 else [False]):
                parts = m.store_gjk_resident(st, wl.pa, st, wl.pb, allgather=gather)
                cover = 0
                for slot, dd, ds, lo, hi in parts:
                    assert (lo, hi) == m.bounds(n, slot)
                    a, b = (0, n) if gather else (lo, hi)
                    dist = np.empty(b - a, dtype=dt)
                    simp = np.empty((b - a) * api.SIMPLEX_F32.itemsize, dtype=np.uint8)
                    rt = _cudart()
                    assert rt.cudaSetDevice(devs[slot]) == 0
                    assert rt.cudaMemcpy(C.c_void_p(dist.ctypes.data), C.c_void_p(dd + a * 4), C.c_size_t(dist.nbytes), 2) == 0
                    assert rt.cudaMemcpy(C.c_void_p(simp.ctypes.data), C.c_void_p(ds + a * api.SIMPLEX_F32.itemsize),
                                         C.c_size_t(simp.nbytes), 2) == 0
                    assert dist.tobytes() == dref[a:b].tobytes(), (devs, slot, gather)
                    got = simp.view(api.SIMPLEX_F32)
                    for f in ("nvrtx", "vrtx", "vrtx_idx", "witnesses"):
                        assert np.ascontiguousarray(got[f]).tobytes() == np.ascontiguousarray(sref[a:b][f]).tobytes(), f
                    cover += hi - lo
                assert cover == n
            st.close()
        finally:
            m.close()


def test_multi_argument_errors():
    with pytest.raises(api.OpenGJKError):
        api.MultiEngine([0, 0])
    with pytest.raises(api.OpenGJKError):
        api.MultiEngine([_ndev() + 3])
    m = api.MultiEngine([0])
    try:
        w = W.mixed_hulls(100, seed=96, dtype=np.float32)
        bad = w.off.copy()
        bad[4] = 10**9
        with pytest.raises(api.OpenGJKError):
            m.gjk_host_pools(w.coords, bad[w.pa], w.cnt[w.pa], w.coords, w.off[w.pb], w.cnt[w.pb])
    finally:
        m.close()
