"""ctypes loader for the CPU oracle and the compiled reference (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  Nothing under opengjk_b200/ does.

Two back ends with one calling convention (flat pool + pair index lists):
  * ``Oracle(prec)``     -- oracle/gjk_oracle.c, our restatement of
    /root/reference/scalar/openGJK.c:941-1046.
  * ``Reference(prec)``  -- the unmodified reference compiled into
    oracle/_ref/libopengjk_ref_{f64,f32}.so (built by oracle/Makefile while
    /root/reference is mounted; the .so travels to the GPU box).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _real(prec):
    return {"f64": (np.float64, C.c_double), "f32": (np.float32, C.c_float)}[prec]


def simplex_dtype(prec):
    """numpy mirror of gkSimplex (scalar/include/openGJK/openGJK.h:84-94)."""
    npf, _ = _real(prec)
    return np.dtype(
        [("nvrtx", np.int32), ("vrtx", npf, (4, 3)), ("vrtx_idx", np.int32, (4, 2)),
         ("witnesses", npf, (2, 3))], align=True)


def build(ref=True):
    """(Re)build the oracle libraries with oracle/Makefile."""
    subprocess.run(["make", "-s", "-C", HERE], check=True)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class _Base:
    def __init__(self, prec, path):
        self.prec = prec
        self.npf, self.cf = _real(prec)
        self.sdt = simplex_dtype(prec)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)

    def _prep(self, coords, off, cnt, pa, pb):
        coords = np.ascontiguousarray(coords, dtype=self.npf).reshape(-1, 3)
        off = np.ascontiguousarray(off, dtype=np.int64)
        cnt = np.ascontiguousarray(cnt, dtype=np.int32)
        pa = np.ascontiguousarray(pa, dtype=np.int32)
        pb = np.ascontiguousarray(pb, dtype=np.int32)
        n = pa.shape[0]
        simp = np.zeros(n, dtype=self.sdt)
        dist = np.zeros(n, dtype=self.npf)
        return coords, off, cnt, pa, pb, n, simp, dist


class Oracle(_Base):
    def __init__(self, prec="f64"):
        super().__init__(prec, os.path.join(HERE, f"libgjk_oracle_{prec}.so"))
        sfx = "_" + prec
        self._gjk = getattr(self.lib, "oracle_gjk" + sfx)
        self._gjk.restype = self.cf
        self._batch = getattr(self.lib, "oracle_gjk_batch" + sfx)
        self._batch.restype = None
        self._sub = getattr(self.lib, "oracle_sub" + sfx)
        self._sub.restype = None
        self._h = [getattr(self.lib, f"oracle_hff{i}{sfx}") for i in (1, 2, 3)]
        for h in self._h:
            h.restype = C.c_int
        sz = getattr(self.lib, "oracle_sizeof_simplex" + sfx)
        sz.restype = C.c_int
        assert sz() == self.sdt.itemsize, (sz(), self.sdt.itemsize)

    def gjk(self, a, b):
        a = np.ascontiguousarray(a, dtype=self.npf).reshape(-1, 3)
        b = np.ascontiguousarray(b, dtype=self.npf).reshape(-1, 3)
        s = np.zeros(1, dtype=self.sdt)
        it = C.c_int(0)
        d = self._gjk(C.c_int(a.shape[0]), _ptr(a), C.c_int(b.shape[0]), _ptr(b), _ptr(s),
                      C.byref(it))
        return float(d), s[0], it.value

    def batch(self, coords, off, cnt, pa, pb, nthreads=1, want_iters=False):
        coords, off, cnt, pa, pb, n, simp, dist = self._prep(coords, off, cnt, pa, pb)
        iters = np.zeros(n, dtype=np.int32)
        self._batch(C.c_int64(n), _ptr(coords), _ptr(off), _ptr(cnt), _ptr(pa), _ptr(pb),
                    _ptr(simp), _ptr(dist), _ptr(iters), C.c_int(nthreads))
        return (dist, simp, iters) if want_iters else (dist, simp)

    def gjk_epa_batch(self, coords, off, cnt, pa, pb, nthreads=1, simplices=None, distances=None):
        """GJK + EPA (oracle/epa_oracle.c).  With simplices/distances given, EPA only
        (the reference's compute_epa).  Returns dict(distances, simplices, witness1,
        witness2, normals, epa_iters)."""
        coords, off, cnt, pa, pb, n, simp, dist = self._prep(coords, off, cnt, pa, pb)
        run_gjk = 1
        if simplices is not None:
            simp = np.ascontiguousarray(simplices, dtype=self.sdt).copy()
            dist = np.ascontiguousarray(distances, dtype=self.npf).copy()
            run_gjk = 0
        w1 = np.zeros((n, 3), dtype=self.npf)
        w2 = np.zeros((n, 3), dtype=self.npf)
        nr = np.zeros((n, 3), dtype=self.npf)
        it = np.zeros(n, dtype=np.int32)
        fn = getattr(self.lib, "oracle_gjk_epa_batch_" + self.prec)
        fn.restype = None
        fn(C.c_int64(n), _ptr(coords), _ptr(off), _ptr(cnt), _ptr(pa), _ptr(pb), _ptr(simp), _ptr(dist), _ptr(w1),
           _ptr(w2), _ptr(nr), _ptr(it), C.c_int(run_gjk), C.c_int(nthreads))
        return {"distances": dist, "simplices": simp, "witness1": w1, "witness2": w2, "normals": nr, "epa_iters": it}

    def sub(self, vrtx, idx=None):
        """Run S1D/S2D/S3D on a simplex given as (nvrtx,3); returns (v, simplex)."""
        vrtx = np.asarray(vrtx, dtype=self.npf).reshape(-1, 3)
        s = np.zeros(1, dtype=self.sdt)
        s["nvrtx"] = vrtx.shape[0]
        s["vrtx"][0, : vrtx.shape[0]] = vrtx
        if idx is not None:
            s["vrtx_idx"][0, : vrtx.shape[0]] = idx
        v = np.ones(3, dtype=self.npf)
        self._sub(_ptr(s), _ptr(v))
        return v, s[0]

    def hff(self, which, *pts):
        arrs = [np.ascontiguousarray(p, dtype=self.npf) for p in pts]
        return int(self._h[which - 1](*[_ptr(a) for a in arrs]))


class Reference(_Base):
    def __init__(self, prec="f64"):
        super().__init__(prec, os.path.join(HERE, "_ref", f"libopengjk_ref_{prec}.so"))
        L = self.lib
        L.ref_rows_create.restype = C.c_void_p
        L.ref_rows_create.argtypes = [C.c_void_p, C.c_int64]
        L.ref_rows_free.argtypes = [C.c_void_p]
        L.ref_gjk_single.restype = self.cf
        L.ref_gjk_batch.restype = None
        L.ref_sub.restype = None
        L.ref_sizeof_simplex.restype = C.c_int
        assert L.ref_sizeof_simplex() == self.sdt.itemsize

    @staticmethod
    def available(prec="f64"):
        return os.path.exists(os.path.join(HERE, "_ref", f"libopengjk_ref_{prec}.so"))

    def gjk(self, a, b):
        a = np.ascontiguousarray(a, dtype=self.npf).reshape(-1, 3)
        b = np.ascontiguousarray(b, dtype=self.npf).reshape(-1, 3)
        s = np.zeros(1, dtype=self.sdt)
        d = self.lib.ref_gjk_single(C.c_int(a.shape[0]), _ptr(a), C.c_int(b.shape[0]), _ptr(b),
                                    _ptr(s))
        return float(d), s[0]

    def batch(self, coords, off, cnt, pa, pb, nthreads=1):
        coords, off, cnt, pa, pb, n, simp, dist = self._prep(coords, off, cnt, pa, pb)
        rows = self.lib.ref_rows_create(_ptr(coords), C.c_int64(coords.shape[0]))
        try:
            self.lib.ref_gjk_batch(C.c_int64(n), C.c_void_p(rows), _ptr(off), _ptr(cnt), _ptr(pa),
                                   _ptr(pb), _ptr(simp), _ptr(dist), C.c_int(nthreads))
        finally:
            self.lib.ref_rows_free(C.c_void_p(rows))
        return dist, simp

    class Prepared:
        """Row pointers built once so a timed region holds only the queries."""

        def __init__(self, ref, coords, off, cnt, pa, pb):
            self.ref = ref
            (self.coords, self.off, self.cnt, self.pa, self.pb, self.n, self.simp,
             self.dist) = ref._prep(coords, off, cnt, pa, pb)
            self.rows = ref.lib.ref_rows_create(_ptr(self.coords), C.c_int64(self.coords.shape[0]))

        def run(self, nthreads):
            self.ref.lib.ref_gjk_batch(C.c_int64(self.n), C.c_void_p(self.rows), _ptr(self.off),
                                       _ptr(self.cnt), _ptr(self.pa), _ptr(self.pb),
                                       _ptr(self.simp), _ptr(self.dist), C.c_int(nthreads))
            return self.dist, self.simp

        def close(self):
            if self.rows:
                self.ref.lib.ref_rows_free(C.c_void_p(self.rows))
                self.rows = None

    def sub(self, vrtx, idx=None):
        vrtx = np.asarray(vrtx, dtype=self.npf).reshape(-1, 3)
        s = np.zeros(1, dtype=self.sdt)
        s["nvrtx"] = vrtx.shape[0]
        s["vrtx"][0, : vrtx.shape[0]] = vrtx
        if idx is not None:
            s["vrtx_idx"][0, : vrtx.shape[0]] = idx
        v = np.ones(3, dtype=self.npf)
        self.lib.ref_sub(_ptr(s), _ptr(v))
        return v, s[0]


def polytope_dtype(prec):
    """numpy mirror of the GPU flavour of gkPolytope (gpu/include/openGJK_GPU.h:71-76)."""
    if prec == "f32":
        return np.dtype({"names": ["numpoints", "s", "s_idx", "coord"], "formats": [np.int32, (np.float32, 3), np.int32, np.uint64],
                         "offsets": [0, 4, 16, 24], "itemsize": 32})
    return np.dtype({"names": ["numpoints", "s", "s_idx", "coord"], "formats": [np.int32, (np.float64, 3), np.int32, np.uint64],
                     "offsets": [0, 8, 32, 40], "itemsize": 48})


class ReferenceGPU(_Base):
    """The reference's own CUDA library (gpu/opengjk_gpu.cu compiled unmodified, default
    nvcc flags, by `make -C oracle ref_gpu`): the existing Blackwell-runnable kernels.
    Used as a loose cross-check (it contracts FMAs, so it is not bit-comparable) and as
    the "reference GPU kernel" baseline."""

    def __init__(self, prec="f32", variant=""):
        """variant "" = the reference's default flags; "_nofma" = the same source with -fmad=false."""
        super().__init__(prec, os.path.join(HERE, "_ref", f"libopenGJK_GPU_ref_{prec}{variant}.so"))
        self.pdt = polytope_dtype(prec)

    @staticmethod
    def available(prec="f32", variant=""):
        return os.path.exists(os.path.join(HERE, "_ref", f"libopenGJK_GPU_ref_{prec}{variant}.so"))

    def _bodies(self, coords, off, cnt, which):
        arr = np.zeros(which.shape[0], dtype=self.pdt)
        arr["numpoints"] = cnt[which]
        arr["coord"] = coords.ctypes.data + off[which].astype(np.uint64) * np.uint64(3 * coords.itemsize)
        return arr

    def gjk(self, coords, off, cnt, pa, pb, epa=False):
        coords, off, cnt, pa, pb, n, simp, dist = self._prep(coords, off, cnt, pa, pb)
        b1, b2 = self._bodies(coords, off, cnt, pa), self._bodies(coords, off, cnt, pb)
        if not epa:
            self.lib.compute_minimum_distance(C.c_int(n), _ptr(b1), _ptr(b2), _ptr(simp), _ptr(dist))
            return dist, simp
        w1 = np.zeros((n, 3), dtype=self.npf)
        w2 = np.zeros((n, 3), dtype=self.npf)
        self.lib.compute_gjk_epa(C.c_int(n), _ptr(b1), _ptr(b2), _ptr(simp), _ptr(dist), _ptr(w1), _ptr(w2))
        return dist, simp, w1, w2

    def epa(self, coords, off, cnt, pa, pb, simplices, distances):
        """The reference's compute_epa (gpu/include/openGJK_GPU.h:176-185): EPA only, on the GJK
        results given.  Returns (distances, witness1, witness2, contact_normals)."""
        coords, off, cnt, pa, pb, n, _, _ = self._prep(coords, off, cnt, pa, pb)
        b1, b2 = self._bodies(coords, off, cnt, pa), self._bodies(coords, off, cnt, pb)
        simp = np.ascontiguousarray(simplices, dtype=self.sdt).copy()
        dist = np.ascontiguousarray(distances, dtype=self.npf).copy()
        w1, w2, nr = (np.zeros((n, 3), dtype=self.npf) for _ in range(3))
        self.lib.compute_epa(C.c_int(n), _ptr(b1), _ptr(b2), _ptr(simp), _ptr(dist), _ptr(w1), _ptr(w2), _ptr(nr))
        return dist, w1, w2, nr

    def time_kernels(self, coords, off, cnt, pa, pb, reps=5, epa=False):
        """Kernel-only milliseconds through the mid-level API (device arrays built once;
        compute_*_device synchronises internally, gpu/opengjk_gpu.cu:3198-3201)."""
        import time
        coords, off, cnt, pa, pb, n, simp, dist = self._prep(coords, off, cnt, pa, pb)
        b1, b2 = self._bodies(coords, off, cnt, pa), self._bodies(coords, off, cnt, pb)
        vp = C.c_void_p
        d = [vp() for _ in range(6)]
        L = self.lib
        L.allocate_and_copy_device_arrays(C.c_int(n), _ptr(b1), _ptr(b2), *[C.byref(x) for x in d])
        e = [vp() for _ in range(3)]
        if epa:
            L.allocate_epa_device_arrays(C.c_int(n), *[C.byref(x) for x in e])
        out = {}
        L.compute_minimum_distance_device(C.c_int(n), d[0], d[1], d[4], d[5])
        t0 = time.perf_counter()
        for _ in range(reps):
            L.compute_minimum_distance_device(C.c_int(n), d[0], d[1], d[4], d[5])
        out["gjk_ms"] = 1e3 * (time.perf_counter() - t0) / reps
        if epa:
            L.compute_epa_device(C.c_int(n), d[0], d[1], d[4], d[5], e[0], e[1], e[2])
            t0 = time.perf_counter()
            for _ in range(reps):
                L.compute_minimum_distance_device(C.c_int(n), d[0], d[1], d[4], d[5])
                L.compute_epa_device(C.c_int(n), d[0], d[1], d[4], d[5], e[0], e[1], e[2])
            out["gjk_epa_ms"] = 1e3 * (time.perf_counter() - t0) / reps
            L.free_epa_device_arrays(e[0], e[1], e[2])
        L.free_device_arrays(*d)
        return out
