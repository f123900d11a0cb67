"""Synthetic GJK workloads (host side, numpy) -- SURVEY.md section 8(d).

A workload is a polytope *pool* plus a *pair list*:

    coords : (nverts, 3) float   all vertices of all polytopes, xyz interleaved
    off    : (npoly,)  int64     first vertex of each polytope
    cnt    : (npoly,)  int32     vertex count of each polytope
    pa, pb : (npairs,) int32     polytope index of body 1 / body 2

which is the flat form of the reference's array-of-gkPolytope inputs
(gpu/include/openGJK_GPU.h:71-76, one contiguous `coord` block per polytope)
and of its indexed mode (gkCollisionPair, openGJK_GPU.h:89-92).
Coordinates are computed in float64 and rounded once to the requested dtype.
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class Workload:
    name: str
    coords: np.ndarray
    off: np.ndarray
    cnt: np.ndarray
    pa: np.ndarray
    pb: np.ndarray

    @property
    def npairs(self):
        return int(self.pa.shape[0])

    @property
    def npoly(self):
        return int(self.cnt.shape[0])

    def astype(self, dtype):
        return Workload(self.name, self.coords.astype(dtype), self.off, self.cnt, self.pa, self.pb)

    def algorithmic_bytes(self):
        """Compulsory bytes per launch: one read of both bodies of every pair
        plus one gkSimplex and one distance written per pair."""
        t = self.coords.dtype.itemsize
        simplex = 184 if t == 8 else 108
        verts = int(self.cnt[self.pa].astype(np.int64).sum() + self.cnt[self.pb].astype(np.int64).sum())
        return verts * 3 * t + self.npairs * (simplex + t)

    def polytope(self, h):
        o = int(self.off[h])
        return self.coords[o:o + int(self.cnt[h])]

    def two_pools(self):
        """(coords1, off1, cnt1, coords2, off2, cnt2) with pair p = polytope p of each pool:
        the un-indexed bd1[] / bd2[] form of the reference's batched API."""
        out = []
        for idx in (self.pa, self.pb):
            cnt = np.ascontiguousarray(self.cnt[idx])
            off = _offsets(cnt)
            start = self.off[idx]
            src = np.repeat(start - off, cnt) + np.arange(int(cnt.sum()), dtype=np.int64)
            out += [np.ascontiguousarray(self.coords[src]), off, cnt]
        return tuple(out)

    def slice_pairs(self, lo, hi):
        return Workload(self.name, self.coords, self.off, self.cnt, self.pa[lo:hi], self.pb[lo:hi])


def _offsets(cnt):
    off = np.zeros(cnt.shape[0], dtype=np.int64)
    np.cumsum(cnt[:-1], out=off[1:])
    return off


def _sphere_points(rng, n, radius=1.0):
    z = 2.0 * rng.random(n) - 1.0
    th = 2.0 * np.pi * rng.random(n)
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    p = np.stack([r * np.cos(th), r * np.sin(th), z], axis=1)
    return p * np.asarray(radius).reshape(-1, 1) if np.ndim(radius) else p * radius


def random_hulls(npairs, nmin=8, nmax=64, seed=42, center_scale=4.0, dtype=np.float64,
                 radius_jitter=0.0):
    """Config #2 / #5: every pair owns two hulls with U{nmin..nmax} vertices on a
    sphere; body 1 at the origin, body 2 centred at center_scale*(u-0.5) per axis.
    center_scale may be an array of per-pair scales (mixed slices)."""
    rng = np.random.default_rng(seed)
    cnt = rng.integers(nmin, nmax + 1, size=2 * npairs).astype(np.int32)
    off = _offsets(cnt)
    total = int(cnt.sum())
    if radius_jitter:
        rad = np.repeat(1.0 + radius_jitter * rng.random(2 * npairs), cnt)
        pts = _sphere_points(rng, total, rad)
    else:
        pts = _sphere_points(rng, total)
    scale = np.broadcast_to(np.asarray(center_scale, dtype=np.float64), (npairs,))
    centre = np.zeros((2 * npairs, 3))
    centre[1::2] = scale[:, None] * (rng.random((npairs, 3)) - 0.5)
    pts += np.repeat(centre, cnt, axis=0)
    pa = np.arange(0, 2 * npairs, 2, dtype=np.int32)
    pb = pa + 1
    return Workload(f"hulls{nmin}-{nmax}", pts.astype(dtype), off, cnt, pa, pb)


def mixed_hulls(npairs, nmin=8, nmax=64, seed=42, dtype=np.float64):
    """Config #2 as specified in SURVEY 8(d): 80 % centre scale 4 (about 38 %
    intersecting), 10 % scale 1 (all intersecting), 10 % scale 10 (all separated)."""
    scale = np.full(npairs, 4.0)
    scale[int(0.8 * npairs):int(0.9 * npairs)] = 1.0
    scale[int(0.9 * npairs):] = 10.0
    w = random_hulls(npairs, nmin, nmax, seed, scale, dtype)
    w.name = f"mixed_hulls{nmin}-{nmax}"
    return w


_BOX = np.array([[sx, sy, sz] for sz in (-1, 1) for sy in (-1, 1) for sx in (-1, 1)],
                dtype=np.float64)  # x fastest, as gpu/examples/python/test_examples.py:275-278


def _random_rotations(rng, n):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    return np.stack([
        np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], axis=1),
        np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], axis=1),
        np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], axis=1),
    ], axis=1)


def boxes(npairs, oriented=True, seed=42, offset_scale=6.0, dtype=np.float32):
    """Config #3: 8-vertex boxes.  oriented=False gives axis-aligned unit cubes
    (the tie-stress case); oriented=True gives half-extents U[0.5,1.5] under a
    uniform random rotation.  Body 2 is offset by offset_scale*(u-0.5) per axis."""
    rng = np.random.default_rng(seed)
    nb = 2 * npairs
    if oriented:
        half = 0.5 + rng.random((nb, 1, 3))
        rot = _random_rotations(rng, nb)
        pts = np.einsum("bij,bvj->bvi", rot, _BOX[None] * half)
    else:
        pts = np.broadcast_to(_BOX[None], (nb, 8, 3)).copy()
    shift = np.zeros((nb, 1, 3))
    shift[1::2, 0] = offset_scale * (rng.random((npairs, 3)) - 0.5)
    pts = (pts + shift).reshape(-1, 3)
    cnt = np.full(nb, 8, dtype=np.int32)
    off = _offsets(cnt)
    pa = np.arange(0, nb, 2, dtype=np.int32)
    return Workload("boxes_oriented" if oriented else "boxes_aligned", pts.astype(dtype), off, cnt,
                    pa, pa + 1)


def large_hull_pool(npairs, npool=256, nmin=1000, nmax=10000, seed=42, radius=2.0,
                    offset_scale=6.0, dtype=np.float32):
    """Config #4 (indexed mode): a pool of large hulls on a radius-2 sphere;
    first half centred at the origin, second half each pre-translated by its own
    offset; pair = (random hull of half A, random hull of half B)."""
    rng = np.random.default_rng(seed)
    cnt = rng.integers(nmin, nmax + 1, size=npool).astype(np.int32)
    off = _offsets(cnt)
    pts = _sphere_points(rng, int(cnt.sum()), radius)
    centre = np.zeros((npool, 3))
    half = npool // 2
    centre[half:] = offset_scale * (rng.random((npool - half, 3)) - 0.5)
    pts += np.repeat(centre, cnt, axis=0)
    pa = rng.integers(0, half, size=npairs).astype(np.int32)
    pb = rng.integers(half, npool, size=npairs).astype(np.int32)
    return Workload(f"pool{nmin}-{nmax}", pts.astype(dtype), off, cnt, pa, pb)


def tiny_bodies(npairs_per_shape=200, seed=7, scale=1.0, dtype=np.float64):
    """1..7-point bodies in [0,scale)^3 -- the degenerate inputs of the reference's
    test_random_objects / test_large_random_objects
    (scalar/examples/python_ctypes/test/test_min_distance.py:197-223)."""
    rng = np.random.default_rng(seed)
    cnts = []
    for i in range(1, 8):
        for j in range(1, 8):
            cnts.append(np.tile(np.array([i, j], dtype=np.int32), npairs_per_shape))
    cnt = np.concatenate(cnts)
    off = _offsets(cnt)
    pts = scale * rng.random((int(cnt.sum()), 3))
    pa = np.arange(0, cnt.shape[0], 2, dtype=np.int32)
    return Workload(f"tiny_s{scale:g}", pts.astype(dtype), off, cnt, pa, pa + 1)


def cap_stress(npairs, seed=11, dtype=np.float64, base=None):
    """Pairs that end on the iteration cap (scalar/openGJK.c:1036-1042): rigid motions and tiny
    perturbations of one cap-prone configuration (`base` = (a, b) vertex arrays; the tests pass
    tests/golden/cap_pair.npz, a pair of the default bench workload).  Roughly 60-100 % of the
    variants run all 25 iterations in either precision; the mix is rotate / translate /
    jitter 1e-9 / jitter 1e-6 / slide body 2 by ~1e-3 / power-of-two scale, in equal parts."""
    a, b = (np.asarray(x, dtype=np.float64) for x in base)
    rng = np.random.default_rng(seed)
    A, B = [], []
    for i in range(npairs):
        kind = i % 6
        if kind == 0:
            q = rng.normal(size=4)
            q /= np.linalg.norm(q)
            w_, x, y, z = q
            r = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w_), 2 * (x * z + y * w_)],
                          [2 * (x * y + z * w_), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w_)],
                          [2 * (x * z - y * w_), 2 * (y * z + x * w_), 1 - 2 * (x * x + y * y)]])
            A.append(a @ r.T); B.append(b @ r.T)
        elif kind == 1:
            t = rng.normal(size=3) * 3
            A.append(a + t); B.append(b + t)
        elif kind == 2:
            A.append(a + rng.normal(size=a.shape) * 1e-9); B.append(b + rng.normal(size=b.shape) * 1e-9)
        elif kind == 3:
            A.append(a + rng.normal(size=a.shape) * 1e-6); B.append(b + rng.normal(size=b.shape) * 1e-6)
        elif kind == 4:
            A.append(a); B.append(b + rng.normal(size=3) * 1e-3)
        else:
            f = 2.0 ** int(rng.integers(-20, 21))
            A.append(a * f); B.append(b * f)
    return from_bodies(A, B, dtype, "cap_stress")


def from_bodies(bodies_a, bodies_b, dtype=np.float64, name="explicit"):
    """Explicit list of (n_i,3) arrays for each side."""
    allb = []
    for a, b in zip(bodies_a, bodies_b):
        allb.append(np.asarray(a, dtype=np.float64).reshape(-1, 3))
        allb.append(np.asarray(b, dtype=np.float64).reshape(-1, 3))
    cnt = np.array([x.shape[0] for x in allb], dtype=np.int32)
    off = _offsets(cnt)
    pa = np.arange(0, len(allb), 2, dtype=np.int32)
    return Workload(name, np.concatenate(allb).astype(dtype), off, cnt, pa, pa + 1)


# Config #1: the reference's two-body example (scalar/examples/c/userP.dat, userQ.dat;
# README.md:55-59 expects 3.653650).
USER_P = np.array([[0.0, 5.5, 0.0], [2.3, 1.0, -2.0], [8.1, 4.0, 2.4], [4.3, 5.0, 2.2],
                   [2.5, 1.0, 2.3], [7.1, 1.0, 2.4], [1.0, 1.5, 0.3], [3.3, 0.5, 0.3],
                   [6.0, 1.4, 0.2]])
USER_Q = np.array([[-0.0, -5.5, 0.0], [-2.3, -1.0, 2.0], [-8.1, -4.0, -2.4], [-4.3, -5.0, -2.2],
                   [-2.5, -1.0, -2.3], [-7.1, -1.0, -2.4], [-1.0, -1.5, -0.3],
                   [-3.3, -0.5, -0.3], [-6.0, -1.4, -0.2]])
