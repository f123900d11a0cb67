"""Host-side mirror of the reference's batched Python interface.

Function names, argument meaning and result dictionaries follow
gpu/examples/python/src/pyopengjk_gpu/opengjk_gpu.py (compute_minimum_distance
:227-318, compute_minimum_distance_indexed :480-560) so that tests read like
the reference's own; the work is done by the C ABI (include/opengjk_b200.h).
Unlike the reference wrapper, no per-polytope ctypes struct is built in a
Python loop: ragged batches are flattened once with numpy and handed over as
one pool.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import OpenGJKError, SIMPLEX_F32, SIMPLEX_F64, check  # noqa: F401

_PREC = {np.dtype(np.float32): ("f32", SIMPLEX_F32), np.dtype(np.float64): ("f64", SIMPLEX_F64)}

MODE_AUTO, MODE_THREAD, MODE_WARP = 0, 1, 2


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.data_ptr())  # torch tensor


class Engine:
    """One CUDA device (ogjk_ctx).  Raises OpenGJKError when no GPU is present."""

    def __init__(self, device=0, lib_path=None):
        self._L = _lib.lib(lib_path)
        self._h = C.c_void_p()
        self._ck(self._L.ogjk_ctx_create(int(device), C.byref(self._h)))
        self.device = device

    def _ck(self, rc):
        check(rc, self._L)

    def pool_status(self):
        """Synchronises the device; raises if a raw pool failed validation since the last report."""
        self._ck(self._L.ogjk_ctx_pool_status(self._h))

    def close(self):
        if self._h:
            self._L.ogjk_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self):
        return int(self._L.ogjk_ctx_launch_count(self._h))

    def set_timing(self, enable=True):
        self._ck(self._L.ogjk_ctx_set_timing(self._h, int(bool(enable))))

    def last_timing(self):
        """Milliseconds of the last device call: repack, sort, and one entry per size class."""
        ms = (C.c_float * 8)()
        self._ck(self._L.ogjk_ctx_last_timing(self._h, ms))
        return {"repack": ms[0], "sort": ms[1], "classes": [ms[i] for i in range(2, 8)]}

    def set_warp_threshold(self, nverts):
        self._ck(self._L.ogjk_ctx_set_warp_threshold(self._h, int(nverts)))

    def set_classes(self, table):
        """table: list of (hi_key, lanes); see ogjk_ctx_set_classes."""
        n = len(table)
        hi = (C.c_int * n)(*[int(t[0]) for t in table])
        ln = (C.c_int * n)(*[int(t[1]) for t in table])
        self._ck(self._L.ogjk_ctx_set_classes(self._h, n, hi, ln))

    # -- device-resident batch over raw xyz-interleaved pools (torch CUDA tensors or raw pointers)
    def gjk_device(self, prec, npairs, coords1, off1, cnt1, idx1, coords2, off2, cnt2, idx2, simplices, distances,
                   mode=MODE_AUTO, stream=0, npoly1=None, nverts1=None, npoly2=None, nverts2=None):
        def size(t, given, div=1):
            if given is not None:
                return int(given)
            return int(t.numel() if hasattr(t, "numel") else t.size) // div
        fn = getattr(self._L, f"ogjk_gjk_device_{prec}")
        self._ck(fn(self._h, int(npairs), _p(coords1), _p(off1), _p(cnt1), size(cnt1, npoly1), size(coords1, nverts1, 3),
                 _p(idx1), _p(coords2), _p(off2), _p(cnt2), size(cnt2, npoly2), size(coords2, nverts2, 3), _p(idx2),
                 _p(simplices), _p(distances), int(mode), C.c_void_p(int(stream))))

    # -- packed store -------------------------------------------------------------
    def store_create(self, coords, off, cnt, stream=0):
        """coords/off/cnt: numpy arrays (host) or torch CUDA tensors (device)."""
        on_device = not isinstance(coords, np.ndarray)
        if on_device:
            prec = "f64" if coords.element_size() == 8 else "f32"
            npoly, nverts = int(cnt.numel()), int(coords.numel()) // 3
        else:
            coords = np.ascontiguousarray(coords)
            prec = _PREC[coords.dtype][0]
            off = np.ascontiguousarray(off, dtype=np.int64)
            cnt = np.ascontiguousarray(cnt, dtype=np.int32)
            npoly, nverts = int(cnt.shape[0]), int(coords.size) // 3
        h = C.c_void_p()
        fn = getattr(self._L, f"ogjk_store_create_{prec}")
        self._ck(fn(self._h, npoly, nverts, _p(coords), _p(off), _p(cnt), int(on_device), C.c_void_p(int(stream)),
                 C.byref(h)))
        return Store(h, prec, self._L)

    def store_gjk(self, s1, idx1, s2, idx2, npairs, simplices, distances, stream=0):
        fn = getattr(self._L, f"ogjk_store_gjk_{s1.prec}")
        self._ck(fn(self._h, s1._h, _p(idx1), s2._h, _p(idx2), int(npairs), _p(simplices), _p(distances),
                 C.c_void_p(int(stream))))

    # -- host buffers, one pool shared by both bodies (indexed form) -------------
    def gjk_host(self, coords, off, cnt, pa, pb, want_simplices=True, out=None):
        coords = np.ascontiguousarray(coords)
        prec, sdt = _PREC[coords.dtype]
        coords = coords.reshape(-1, 3)
        off = np.ascontiguousarray(off, dtype=np.int64)
        cnt = np.ascontiguousarray(cnt, dtype=np.int32)
        pa = np.ascontiguousarray(pa, dtype=np.int32)
        pb = np.ascontiguousarray(pb, dtype=np.int32)
        n = pa.shape[0]
        if out is None:
            dist = np.zeros(n, dtype=coords.dtype)
            simp = np.zeros(n, dtype=sdt) if want_simplices else None
        else:
            dist, simp = out
        fn = getattr(self._L, f"ogjk_gjk_host_{prec}")
        self._ck(fn(self._h, n, _p(coords), _p(off), _p(cnt), cnt.shape[0], coords.shape[0], _p(pa), _p(coords), _p(off),
                 _p(cnt), cnt.shape[0], coords.shape[0], _p(pb), _p(simp), _p(dist)))
        return dist, simp

    def store_epa(self, s1, idx1, s2, idx2, npairs, simplices, distances, witness1, witness2, normals=None, stream=0):
        fn = getattr(self._L, f"ogjk_store_epa_{s1.prec}")
        self._ck(fn(self._h, s1._h, _p(idx1), s2._h, _p(idx2), int(npairs), _p(simplices), _p(distances), _p(witness1),
                 _p(witness2), _p(normals), C.c_void_p(int(stream))))

    def store_gjk_epa(self, s1, idx1, s2, idx2, npairs, simplices, distances, witness1, witness2, normals=None,
                      stream=0):
        fn = getattr(self._L, f"ogjk_store_gjk_epa_{s1.prec}")
        self._ck(fn(self._h, s1._h, _p(idx1), s2._h, _p(idx2), int(npairs), _p(simplices), _p(distances), _p(witness1),
                 _p(witness2), _p(normals), C.c_void_p(int(stream))))

    def gjk_epa_host(self, coords, off, cnt, pa, pb, want_normals=True):
        """GJK + EPA over host buffers (one shared pool).  Returns a dict of arrays."""
        coords = np.ascontiguousarray(coords)
        prec, sdt = _PREC[coords.dtype]
        coords = coords.reshape(-1, 3)
        off = np.ascontiguousarray(off, dtype=np.int64)
        cnt = np.ascontiguousarray(cnt, dtype=np.int32)
        pa = np.ascontiguousarray(pa, dtype=np.int32)
        pb = np.ascontiguousarray(pb, dtype=np.int32)
        n = pa.shape[0]
        dist = np.zeros(n, dtype=coords.dtype)
        simp = np.zeros(n, dtype=sdt)
        w1 = np.zeros((n, 3), dtype=coords.dtype)
        w2 = np.zeros((n, 3), dtype=coords.dtype)
        nr = np.zeros((n, 3), dtype=coords.dtype) if want_normals else None
        fn = getattr(self._L, f"ogjk_gjk_epa_host_{prec}")
        self._ck(fn(self._h, n, _p(coords), _p(off), _p(cnt), cnt.shape[0], coords.shape[0], _p(pa), _p(coords), _p(off),
                 _p(cnt), cnt.shape[0], coords.shape[0], _p(pb), _p(simp), _p(dist), _p(w1), _p(w2), _p(nr)))
        return {"distances": dist, "simplices": simp, "witness1": w1, "witness2": w2, "normals": nr}

    # -- broad-phase-adjacent pair production (device pointers; include/opengjk_b200.h) --------
    def store_aabbs(self, store, boxes, stream=0):
        """boxes: device buffer of 6*num_polytopes values, filled with {min xyz, max xyz} per polytope."""
        self._ck(getattr(self._L, f"ogjk_store_aabbs_{store.prec}")(self._h, store._h, _p(boxes), C.c_void_p(int(stream))))

    def pairs_cull(self, prec, boxes1, boxes2, ncand, idx1, idx2, margin, capacity, out1, out2, count, stream=0):
        self._ck(getattr(self._L, f"ogjk_pairs_cull_{prec}")(self._h, _p(boxes1), _p(boxes2), int(ncand), _p(idx1), _p(idx2),
                                                            float(margin), int(capacity), _p(out1), _p(out2), _p(count),
                                                            C.c_void_p(int(stream))))

    def pairs_all(self, prec, boxes, npoly, margin, capacity, out1, out2, count, stream=0):
        self._ck(getattr(self._L, f"ogjk_pairs_all_{prec}")(self._h, _p(boxes), int(npoly), float(margin), int(capacity),
                                                           _p(out1), _p(out2), _p(count), C.c_void_p(int(stream))))

    def intersect_flags(self, prec, n, distances, flags, stream=0):
        """flags[i] = !(distances[i] > eps) on the device (uint8; 1 = touching or intersecting)."""
        self._ck(getattr(self._L, f"ogjk_intersect_flags_{prec}")(self._h, int(n), _p(distances), _p(flags), C.c_void_p(int(stream))))

    def set_accel_min(self, min_vertices):
        """Vertex count from which stores created afterwards get a cluster table (0 = off)."""
        self._ck(self._L.ogjk_ctx_set_accel_min(self._h, int(min_vertices)))

    def set_passes(self, passes, iters1=4, iters2=3):
        """Multi-pass scheduling of the small-group classes (1 = one-pass kernel)."""
        self._ck(self._L.ogjk_ctx_set_passes(self._h, int(passes), int(iters1), int(iters2)))

    def set_table_lanes(self, lanes):
        """Kernel choice for the cluster-table pairs (0 default routing, 1/2/3 or 8/16/32: see the header)."""
        self._ck(self._L.ogjk_ctx_set_table_lanes(self._h, int(lanes)))

    def set_regroup(self, iters1, iters2=0):
        """In-CTA regrouping of the 2-lane classes after iters1 (and iters2) iterations; 0 = off."""
        self._ck(self._L.ogjk_ctx_set_regroup(self._h, int(iters1), int(iters2)))

    def set_pipeline_chunks(self, chunks):
        self._ck(self._L.ogjk_ctx_set_pipeline_chunks(self._h, int(chunks)))

    def gjk_host_pools(self, c1, o1, k1, c2, o2, k2, out=None, epa=False, epa_out=None):
        """Un-indexed form over two pools (pair p = polytope p of each pool), the shape of
        the reference's bd1[] / bd2[] arrays; large batches are chunk-pipelined over PCIe.
        out = (distances, simplices), epa_out = (witness1, witness2, normals): caller-owned outputs."""
        c1 = np.ascontiguousarray(c1); c2 = np.ascontiguousarray(c2)
        if c1.dtype != c2.dtype:
            raise ValueError(f"pools differ in precision: {c1.dtype} vs {c2.dtype}")
        prec, sdt = _PREC[c1.dtype]
        o1 = np.ascontiguousarray(o1, dtype=np.int64); o2 = np.ascontiguousarray(o2, dtype=np.int64)
        k1 = np.ascontiguousarray(k1, dtype=np.int32); k2 = np.ascontiguousarray(k2, dtype=np.int32)
        if not (k1.shape[0] == k2.shape[0] == o1.shape[0] == o2.shape[0]):
            raise ValueError("un-indexed form: both pools need one polytope per pair")
        n = int(k1.shape[0])
        if out is None:
            dist = np.zeros(n, dtype=c1.dtype)
            simp = np.zeros(n, dtype=sdt)
        else:
            dist, simp = out
        args = [self._h, n, _p(c1), _p(o1), _p(k1), n, c1.size // 3, None, _p(c2), _p(o2), _p(k2), n, c2.size // 3, None,
                _p(simp), _p(dist)]
        if epa:
            if epa_out is None:
                w1 = np.zeros((n, 3), dtype=c1.dtype); w2 = np.zeros((n, 3), dtype=c1.dtype); nr = np.zeros((n, 3), dtype=c1.dtype)
            else:
                w1, w2, nr = epa_out  # caller-owned (e.g. pinned) output arrays
            self._ck(getattr(self._L, f"ogjk_gjk_epa_host_{prec}")(*args, _p(w1), _p(w2), _p(nr)))
            return {"distances": dist, "simplices": simp, "witness1": w1, "witness2": w2, "normals": nr}
        self._ck(getattr(self._L, f"ogjk_gjk_host_{prec}")(*args))
        return dist, simp

    def run_workload(self, w, want_simplices=True):
        return self.gjk_host(w.coords, w.off, w.cnt, w.pa, w.pb, want_simplices)


class Store:
    """Handle of a packed device-resident polytope pool (ogjk_store)."""

    def __init__(self, handle, prec, L=None):
        self._h, self.prec, self._L = handle, prec, L or _lib.lib()

    @property
    def nbytes(self):
        return int(self._L.ogjk_store_bytes(self._h))

    @property
    def table_bytes(self):
        """Bytes of cluster tables (large polytopes' pruned-scan accelerator); 0 = none."""
        return int(self._L.ogjk_store_table_bytes(self._h))

    @property
    def num_polytopes(self):
        return int(self._L.ogjk_store_num_polytopes(self._h))

    def close(self):
        if self._h:
            self._L.ogjk_store_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def shard_bounds(npairs, ndev, i, lib_path=None):
    """[lo, hi) of the pair list that device slot i of ndev owns (ogjk_multi_shard_bounds: the
    library owns the partition; a pure function, no GPU needed)."""
    lo, hi = C.c_int64(), C.c_int64()
    _lib.lib(lib_path).ogjk_multi_shard_bounds(int(npairs), int(ndev), int(i), C.byref(lo), C.byref(hi))
    return int(lo.value), int(hi.value)


class MultiEngine:
    """Several GPUs of one box behind one call (ogjk_multi): every device owns a contiguous
    slice of the pair list, results land in disjoint ranges of the caller's arrays."""

    def __init__(self, devices=None, lib_path=None):
        self._L = _lib.lib(lib_path)
        self._h = C.c_void_p()
        if devices is None:
            rc = self._L.ogjk_multi_create(None, 0, C.byref(self._h))
        else:
            devices = [int(d) for d in devices]
            arr = (C.c_int * len(devices))(*devices)
            rc = self._L.ogjk_multi_create(C.cast(arr, C.c_void_p), len(devices), C.byref(self._h))
        check(rc, self._L)
        self.ndev = int(self._L.ogjk_multi_num_devices(self._h))

    def _ck(self, rc):
        check(rc, self._L)

    def close(self):
        if self._h:
            self._L.ogjk_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self):
        return int(self._L.ogjk_multi_launch_count(self._h))

    def bounds(self, npairs, i):
        return shard_bounds(npairs, self.ndev, i)

    def compute_minimum_distance_structs(self, bodies1, bodies2, dtype=np.float32, epa=False):
        """The reference's call shape (arrays of gkPolytope pointing at host vertex arrays) over every device:
        ogjk_multi_compute_minimum_distance_* / ogjk_multi_compute_gjk_epa_*."""
        prec, sdt = _PREC[np.dtype(dtype)]
        a1, _k1 = _struct_array(bodies1, dtype)
        a2, _k2 = _struct_array(bodies2, dtype)
        n = len(bodies1)
        simp = np.zeros(n, dtype=sdt)
        dist = np.zeros(n, dtype=dtype)
        if not epa:
            fn = getattr(self._L, f"ogjk_multi_compute_minimum_distance_{prec}")
            self._ck(fn(self._h, n, C.cast(a1, C.c_void_p), C.cast(a2, C.c_void_p), _p(simp), _p(dist)))
            return dist, simp
        w1, w2 = np.zeros((n, 3), dtype=dtype), np.zeros((n, 3), dtype=dtype)
        fn = getattr(self._L, f"ogjk_multi_compute_gjk_epa_{prec}")
        self._ck(fn(self._h, n, C.cast(a1, C.c_void_p), C.cast(a2, C.c_void_p), _p(simp), _p(dist), _p(w1), _p(w2)))
        return {"distances": dist, "simplices": simp, "witness1": w1, "witness2": w2}

    def gjk_host(self, coords, off, cnt, pa, pb, want_simplices=True, out=None):
        """Indexed form over one shared host pool (replicated on every device)."""
        coords = np.ascontiguousarray(coords)
        prec, sdt = _PREC[coords.dtype]
        coords = coords.reshape(-1, 3)
        off = np.ascontiguousarray(off, dtype=np.int64)
        cnt = np.ascontiguousarray(cnt, dtype=np.int32)
        pa = np.ascontiguousarray(pa, dtype=np.int32)
        pb = np.ascontiguousarray(pb, dtype=np.int32)
        n = pa.shape[0]
        if out is None:
            dist = np.zeros(n, dtype=coords.dtype)
            simp = np.zeros(n, dtype=sdt) if want_simplices else None
        else:
            dist, simp = out
        fn = getattr(self._L, f"ogjk_multi_gjk_host_{prec}")
        self._ck(fn(self._h, n, _p(coords), _p(off), _p(cnt), cnt.shape[0], coords.shape[0], _p(pa), _p(coords), _p(off),
                    _p(cnt), cnt.shape[0], coords.shape[0], _p(pb), _p(simp), _p(dist)))
        return dist, simp

    def gjk_host_pools(self, c1, o1, k1, c2, o2, k2, out=None, epa=False):
        """Un-indexed form over two pools (pair p = polytope p of each): every device receives
        only the coordinates of its own slice."""
        c1 = np.ascontiguousarray(c1); c2 = np.ascontiguousarray(c2)
        if c1.dtype != c2.dtype:
            raise ValueError(f"pools differ in precision: {c1.dtype} vs {c2.dtype}")
        prec, sdt = _PREC[c1.dtype]
        o1 = np.ascontiguousarray(o1, dtype=np.int64); o2 = np.ascontiguousarray(o2, dtype=np.int64)
        k1 = np.ascontiguousarray(k1, dtype=np.int32); k2 = np.ascontiguousarray(k2, dtype=np.int32)
        n = int(k1.shape[0])
        if out is None:
            dist = np.zeros(n, dtype=c1.dtype)
            simp = np.zeros(n, dtype=sdt)
        else:
            dist, simp = out
        args = [self._h, n, _p(c1), _p(o1), _p(k1), n, c1.size // 3, None, _p(c2), _p(o2), _p(k2), n, c2.size // 3, None,
                _p(simp), _p(dist)]
        if epa:
            w1 = np.zeros((n, 3), dtype=c1.dtype); w2 = np.zeros((n, 3), dtype=c1.dtype); nr = np.zeros((n, 3), dtype=c1.dtype)
            self._ck(getattr(self._L, f"ogjk_multi_gjk_epa_host_{prec}")(*args, _p(w1), _p(w2), _p(nr)))
            return {"distances": dist, "simplices": simp, "witness1": w1, "witness2": w2, "normals": nr}
        self._ck(getattr(self._L, f"ogjk_multi_gjk_host_{prec}")(*args))
        return dist, simp

    def store_create(self, coords, off, cnt, via_nccl=False):
        coords = np.ascontiguousarray(coords)
        prec = _PREC[coords.dtype][0]
        off = np.ascontiguousarray(off, dtype=np.int64)
        cnt = np.ascontiguousarray(cnt, dtype=np.int32)
        h = C.c_void_p()
        self._ck(getattr(self._L, f"ogjk_multi_store_create_{prec}")(self._h, int(cnt.shape[0]), int(coords.size) // 3,
                                                                    _p(coords), _p(off), _p(cnt), int(bool(via_nccl)),
                                                                    C.byref(h)))
        return MultiStore(h, prec, self._L)

    def store_gjk(self, s1, pa, s2, pb, out=None, epa=False):
        """One batch over replicated stores; pa/pb are HOST index arrays."""
        pa = np.ascontiguousarray(pa, dtype=np.int32)
        pb = np.ascontiguousarray(pb, dtype=np.int32)
        n = int(pa.shape[0])
        ft, sdt = (np.float64, SIMPLEX_F64) if s1.prec == "f64" else (np.float32, SIMPLEX_F32)
        dist, simp = out if out is not None else (np.zeros(n, dtype=ft), np.zeros(n, dtype=sdt))
        if epa:
            w1, w2, nr = (np.zeros((n, 3), dtype=ft) for _ in range(3))
            self._ck(getattr(self._L, f"ogjk_multi_store_gjk_epa_{s1.prec}")(self._h, s1._h, _p(pa), s2._h, _p(pb), n,
                                                                            _p(simp), _p(dist), _p(w1), _p(w2), _p(nr)))
            return {"distances": dist, "simplices": simp, "witness1": w1, "witness2": w2, "normals": nr}
        self._ck(getattr(self._L, f"ogjk_multi_store_gjk_{s1.prec}")(self._h, s1._h, _p(pa), s2._h, _p(pb), n, _p(simp),
                                                                    _p(dist)))
        return dist, simp

    def store_gjk_resident(self, s1, pa, s2, pb, allgather=False):
        """Results stay on the devices; returns [(device slot, d_dist ptr, d_simp ptr, lo, hi)].
        With allgather=True every device ends up with all results (NCCL over NVLink)."""
        pa = np.ascontiguousarray(pa, dtype=np.int32)
        pb = np.ascontiguousarray(pb, dtype=np.int32)
        n = int(pa.shape[0])
        self._ck(getattr(self._L, f"ogjk_multi_store_gjk_resident_{s1.prec}")(self._h, s1._h, _p(pa), s2._h, _p(pb), n))
        if allgather:
            self._ck(getattr(self._L, f"ogjk_multi_allgather_{s1.prec}")(self._h))
        out = []
        for i in range(self.ndev):
            dd, ds, lo, hi = C.c_void_p(), C.c_void_p(), C.c_int64(), C.c_int64()
            self._ck(getattr(self._L, f"ogjk_multi_resident_{s1.prec}")(self._h, i, C.byref(dd), C.byref(ds), C.byref(lo),
                                                                       C.byref(hi)))
            out.append((i, dd.value, ds.value, int(lo.value), int(hi.value)))
        return out


class MultiStore:
    """A pool replicated on every device of a MultiEngine (ogjk_multi_store)."""

    def __init__(self, handle, prec, L):
        self._h, self.prec, self._L = handle, prec, L

    @property
    def nbytes(self):
        return int(self._L.ogjk_multi_store_bytes(self._h))

    def close(self):
        if self._h:
            self._L.ogjk_multi_store_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default = {}


def default_engine(device=0):
    if device not in _default:
        _default[device] = Engine(device)
    return _default[device]


def _flatten(batch, dtype):
    """(P,V,3) array or list of (V_i,3) arrays -> (coords, off, cnt)."""
    if isinstance(batch, np.ndarray) and batch.ndim == 3:
        if batch.shape[2] != 3:
            raise ValueError(f"Expected 3D array (num_poly, num_verts, 3), got shape {batch.shape}")
        p, v = batch.shape[0], batch.shape[1]
        coords = np.ascontiguousarray(batch, dtype=dtype).reshape(-1, 3)
        cnt = np.full(p, v, dtype=np.int32)
        off = np.arange(p, dtype=np.int64) * v
        return coords, off, cnt
    if isinstance(batch, np.ndarray) and batch.ndim == 2:
        batch = [batch]
    arrs = []
    for a in batch:
        a = np.asarray(a, dtype=dtype)
        if a.ndim != 2 or a.shape[1] != 3:
            raise ValueError(f"Vertices must have shape (n, 3), got {a.shape}")
        arrs.append(a)
    cnt = np.array([a.shape[0] for a in arrs], dtype=np.int32)
    off = np.zeros(len(arrs), dtype=np.int64)
    np.cumsum(cnt[:-1], out=off[1:])
    coords = np.concatenate(arrs) if arrs else np.zeros((0, 3), dtype=dtype)
    return np.ascontiguousarray(coords), off, cnt


def _result(dist, simp):
    return {
        "distances": dist,
        "witnesses1": np.ascontiguousarray(simp["witnesses"][:, 0, :]),
        "witnesses2": np.ascontiguousarray(simp["witnesses"][:, 1, :]),
        "is_collision": np.abs(dist) < 1e-6,  # opengjk_gpu.py:308
        "simplex_nvrtx": simp["nvrtx"].copy(),
        "simplices": simp,
    }


def compute_minimum_distance(vertices1, vertices2, dtype=np.float32, engine=None):
    """Batched GJK distance (reference: opengjk_gpu.py:227-318).  vertices1/2:
    (V,3) single pair, (P,V,3) batch, or list of (V_i,3) arrays."""
    eng = engine or default_engine()
    c1, o1, k1 = _flatten(vertices1, dtype)
    c2, o2, k2 = _flatten(vertices2, dtype)
    if k1.shape[0] != k2.shape[0]:
        raise ValueError(f"Mismatch: {k1.shape[0]} polytopes in bd1, {k2.shape[0]} in bd2")
    n = k1.shape[0]
    coords = np.concatenate([c1, c2])
    off = np.concatenate([o1, o2 + c1.shape[0]])
    cnt = np.concatenate([k1, k2])
    pa = np.arange(n, dtype=np.int32)
    dist, simp = eng.gjk_host(coords, off, cnt, pa, pa + n)
    return _result(dist, simp)


def compute_minimum_distance_indexed(polytopes, pairs, dtype=np.float32, engine=None):
    """Indexed batch (reference: opengjk_gpu.py:480-560): `polytopes` is a
    (P,V,3) array or list of (V_i,3) arrays, `pairs` an (N,2) int array."""
    eng = engine or default_engine()
    coords, off, cnt = _flatten(polytopes, dtype)
    pairs = np.asarray(pairs, dtype=np.int32).reshape(-1, 2)
    if pairs.size and (pairs.min() < 0 or pairs.max() >= cnt.shape[0]):
        raise ValueError("pair index out of range")
    dist, simp = eng.gjk_host(coords, off, cnt, pairs[:, 0], pairs[:, 1])
    return _result(dist, simp)


def _pair_pool(vertices1, vertices2, dtype):
    c1, o1, k1 = _flatten(vertices1, dtype)
    c2, o2, k2 = _flatten(vertices2, dtype)
    if k1.shape[0] != k2.shape[0]:
        raise ValueError(f"Mismatch: {k1.shape[0]} polytopes in bd1, {k2.shape[0]} in bd2")
    n = k1.shape[0]
    pa = np.arange(n, dtype=np.int32)
    return np.concatenate([c1, c2]), np.concatenate([o1, o2 + c1.shape[0]]), np.concatenate([k1, k2]), pa, pa + n


def compute_gjk_epa(vertices1, vertices2, dtype=np.float32, engine=None):
    """GJK then EPA for colliding pairs (reference: opengjk_gpu.py:393-465).  Distances
    are negative penetration depths for intersecting pairs."""
    eng = engine or default_engine()
    r = eng.gjk_epa_host(*_pair_pool(vertices1, vertices2, dtype), want_normals=False)
    return {"distances": r["distances"], "is_collision": np.abs(r["distances"]) < 1e-6, "witnesses1": r["witness1"],
            "witnesses2": r["witness2"], "simplex_nvrtx": r["simplices"]["nvrtx"].copy(), "simplices": r["simplices"]}


def compute_epa(vertices1, vertices2, return_normals=False, dtype=np.float32, engine=None):
    """Penetration depth, contact points and (optionally) contact normals
    (reference: opengjk_gpu.py:319-390).  Unlike the reference wrapper, which feeds
    the EPA kernel zero-initialised simplices, GJK is run first."""
    eng = engine or default_engine()
    r = eng.gjk_epa_host(*_pair_pool(vertices1, vertices2, dtype), want_normals=return_normals)
    out = {"penetration_depths": r["distances"], "witnesses1": r["witness1"], "witnesses2": r["witness2"]}
    if return_normals:
        out["contact_normals"] = r["normals"]
    return out


def compute_gjk_epa_structs(bodies1, bodies2, dtype=np.float32, engine=None):
    """ogjk_compute_gjk_epa_* followed by ogjk_compute_epa_* semantics check helper:
    returns (dist, simp, w1, w2) of the fused call."""
    eng = engine or default_engine()
    prec, sdt = _PREC[np.dtype(dtype)]
    a1, k1 = _struct_array(bodies1, dtype)
    a2, k2 = _struct_array(bodies2, dtype)
    n = len(bodies1)
    simp = np.zeros(n, dtype=sdt)
    dist = np.zeros(n, dtype=dtype)
    w1 = np.zeros((n, 3), dtype=dtype)
    w2 = np.zeros((n, 3), dtype=dtype)
    fn = getattr(eng._L, f"ogjk_compute_gjk_epa_{prec}")
    eng._ck(fn(eng._h, n, C.cast(a1, C.c_void_p), C.cast(a2, C.c_void_p), _p(simp), _p(dist), _p(w1), _p(w2)))
    return dist, simp, w1, w2


def compute_epa_structs(bodies1, bodies2, simplices, distances, dtype=np.float32, engine=None):
    """ogjk_compute_epa_*: EPA on caller-provided GJK results (simplices, distances)."""
    eng = engine or default_engine()
    prec, sdt = _PREC[np.dtype(dtype)]
    a1, k1 = _struct_array(bodies1, dtype)
    a2, k2 = _struct_array(bodies2, dtype)
    n = len(bodies1)
    simp = np.ascontiguousarray(simplices, dtype=sdt).copy()
    dist = np.ascontiguousarray(distances, dtype=dtype).copy()
    w1 = np.zeros((n, 3), dtype=dtype)
    w2 = np.zeros((n, 3), dtype=dtype)
    nr = np.zeros((n, 3), dtype=dtype)
    fn = getattr(eng._L, f"ogjk_compute_epa_{prec}")
    eng._ck(fn(eng._h, n, C.cast(a1, C.c_void_p), C.cast(a2, C.c_void_p), _p(simp), _p(dist), _p(w1), _p(w2), _p(nr)))
    return dist, w1, w2, nr


# ---- struct-array entry points (the exact reference C signatures) -------------
def _polytope_struct(cf):
    class gkPolytope(C.Structure):  # gpu/include/openGJK_GPU.h:71-76
        _fields_ = [("numpoints", C.c_int), ("s", cf * 3), ("s_idx", C.c_int), ("coord", C.POINTER(cf))]
    return gkPolytope


class gkCollisionPair(C.Structure):  # gpu/include/openGJK_GPU.h:89-92
    _fields_ = [("idx1", C.c_int), ("idx2", C.c_int)]


def _struct_array(bodies, dtype):
    cf = C.c_float if np.dtype(dtype) == np.float32 else C.c_double
    P = _polytope_struct(cf)
    arr = (P * len(bodies))()
    keep = []
    for i, b in enumerate(bodies):
        flat = np.ascontiguousarray(np.asarray(b, dtype=dtype).reshape(-1))
        keep.append(flat)
        arr[i].numpoints = flat.shape[0] // 3
        arr[i].s_idx = 0
        arr[i].coord = flat.ctypes.data_as(C.POINTER(cf))
    return arr, keep


def compute_minimum_distance_structs(bodies1, bodies2, dtype=np.float32, engine=None):
    """Calls ogjk_compute_minimum_distance_{f32,f64} with arrays of gkPolytope."""
    eng = engine or default_engine()
    prec, sdt = _PREC[np.dtype(dtype)]
    a1, k1 = _struct_array(bodies1, dtype)
    a2, k2 = _struct_array(bodies2, dtype)
    n = len(bodies1)
    simp = np.zeros(n, dtype=sdt)
    dist = np.zeros(n, dtype=dtype)
    fn = getattr(eng._L, f"ogjk_compute_minimum_distance_{prec}")
    eng._ck(fn(eng._h, n, C.cast(a1, C.c_void_p), C.cast(a2, C.c_void_p), _p(simp), _p(dist)))
    return dist, simp


def compute_minimum_distance_indexed_structs(bodies, pairs, dtype=np.float32, engine=None):
    eng = engine or default_engine()
    prec, sdt = _PREC[np.dtype(dtype)]
    a, keep = _struct_array(bodies, dtype)
    pairs = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
    n = pairs.shape[0]
    simp = np.zeros(n, dtype=sdt)
    dist = np.zeros(n, dtype=dtype)
    fn = getattr(eng._L, f"ogjk_compute_minimum_distance_indexed_{prec}")
    eng._ck(fn(eng._h, len(bodies), n, C.cast(a, C.c_void_p), _p(pairs), _p(simp), _p(dist)))
    return dist, simp


def subalgorithm(vrtx, idx=None, dtype=np.float64, engine=None):
    """Device sub-solver on one simplex given as (nvrtx,3), newest vertex last
    (reference test hooks opengjk_test_S{1,2,3}D).  Returns (v, simplex)."""
    eng = engine or default_engine()
    prec, sdt = _PREC[np.dtype(dtype)]
    vrtx = np.asarray(vrtx, dtype=dtype).reshape(-1, 3)
    s = np.zeros(1, dtype=sdt)
    s["nvrtx"] = vrtx.shape[0]
    s["vrtx"][0, : vrtx.shape[0]] = vrtx
    if idx is not None:
        s["vrtx_idx"][0, : vrtx.shape[0]] = idx
    v = np.ones(3, dtype=dtype)
    eng._ck(getattr(eng._L, f"ogjk_test_subalgorithm_{prec}")(eng._h, _p(s), _p(v)))
    return v, s[0]


def single_query(vertices0, vertices1, dtype=np.float64, engine=None):
    """One pair through the scalar library's row-pointer layout
    (scalar/examples/python_ctypes/src/pyopengjk/opengjk.py:60-91)."""
    eng = engine or default_engine()
    prec, sdt = _PREC[np.dtype(dtype)]
    cf = C.c_float if prec == "f32" else C.c_double
    a = np.ascontiguousarray(np.asarray(vertices0, dtype=dtype).reshape(-1, 3))
    b = np.ascontiguousarray(np.asarray(vertices1, dtype=dtype).reshape(-1, 3))
    rows_a = (C.POINTER(cf) * a.shape[0])(*[a[i].ctypes.data_as(C.POINTER(cf)) for i in range(a.shape[0])])
    rows_b = (C.POINTER(cf) * b.shape[0])(*[b[i].ctypes.data_as(C.POINTER(cf)) for i in range(b.shape[0])])
    simp = np.zeros(1, dtype=sdt)
    dist = np.zeros(1, dtype=dtype)
    fn = getattr(eng._L, f"ogjk_single_rows_{prec}")
    eng._ck(fn(eng._h, a.shape[0], C.cast(rows_a, C.c_void_p), b.shape[0], C.cast(rows_b, C.c_void_p), _p(simp),
             _p(dist)))
    return float(dist[0]), simp[0]
