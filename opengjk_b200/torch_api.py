"""Zero-copy batched front door for GPU-resident callers.

The reference's Python wrapper builds one ctypes gkPolytope per body in a Python loop
and copies every body to the device separately (gpu/examples/python/src/pyopengjk_gpu/
opengjk_gpu.py:195-220, 302-305); that loop dominates wall time long before the GPU does.
Here a batch is a (P, V, 3) CUDA tensor (or any object exporting __dlpack__), optionally
with per-body true vertex counts for ragged batches; nothing is copied or synchronised:
the tensors' memory goes straight to the C ABI (ogjk_gjk_device_* / ogjk_store_*) on
torch's current stream, and results come back as SoA torch tensors in stream order.

Result keys follow the reference's dictionaries (opengjk_gpu.py:306-318): distances,
witnesses1, witnesses2, is_collision, simplex_nvrtx; plus simplex_vrtx, simplex_vrtx_idx
and the raw gkSimplex bytes.  torch is plumbing here (memory + streams), not compute.
"""
import numpy as np

from . import api
from ._lib import SIMPLEX_F32, SIMPLEX_F64


def _torch():
    import torch
    return torch


def _as_bodies(x):
    """(P, V, 3) contiguous CUDA tensor of fp32/fp64 from a tensor or a DLPack exporter."""
    torch = _torch()
    if not isinstance(x, torch.Tensor):
        x = torch.from_dlpack(x)
    if not x.is_cuda:
        raise ValueError("CUDA tensors required (host arrays go through opengjk_b200.api)")
    if x.dim() == 2:
        x = x.unsqueeze(0)
    if x.dim() != 3 or x.shape[-1] != 3:
        raise ValueError(f"expected (P, V, 3) vertices, got {tuple(x.shape)}")
    if x.dtype not in (torch.float32, torch.float64):
        raise ValueError("vertices must be float32 or float64")
    return x.contiguous()


def _layout(bodies, counts):
    """Flat pool description of a padded (P, V, 3) batch: coords view, int64 offsets, int32 counts."""
    torch = _torch()
    P, V, _ = bodies.shape
    off = torch.arange(P, dtype=torch.int64, device=bodies.device) * V
    if counts is None:
        cnt = torch.full((P,), V, dtype=torch.int32, device=bodies.device)
    else:
        cnt = torch.as_tensor(counts, device=bodies.device).to(torch.int32).contiguous()
        if cnt.shape != (P,):
            raise ValueError(f"counts must have shape ({P},)")
    return bodies.view(-1), off, cnt


def _soa(simp_bytes, dist, n, f64):
    """Split the gkSimplex AoS (scalar/include/openGJK/openGJK.h:84-94) into SoA tensors."""
    torch = _torch()
    sdt = SIMPLEX_F64 if f64 else SIMPLEX_F32
    ft = torch.float64 if f64 else torch.float32
    rows = simp_bytes.view(n, sdt.itemsize)

    def field(name, dtype, shape):
        o = sdt.fields[name][1]
        nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        return rows[:, o:o + nbytes].contiguous().view(dtype).view(n, *shape)

    wit = field("witnesses", ft, (2, 3))
    eps = torch.finfo(ft).eps
    return {
        "distances": dist,
        "witnesses1": wit[:, 0, :],
        "witnesses2": wit[:, 1, :],
        "is_collision": dist.abs() < 1e-6,            # opengjk_gpu.py:308
        "intersecting": ~(dist > eps),                # the reference's EPA gate (gpu/opengjk_gpu.cu:2053)
        "simplex_nvrtx": field("nvrtx", torch.int32, (1,)).view(n),
        "simplex_vrtx": field("vrtx", ft, (4, 3)),
        "simplex_vrtx_idx": field("vrtx_idx", torch.int32, (4, 2)),
        "simplices": simp_bytes,
    }


def _outputs(n, like):
    torch = _torch()
    f64 = like.dtype == torch.float64
    sdt = SIMPLEX_F64 if f64 else SIMPLEX_F32
    simp = torch.empty(n * sdt.itemsize, dtype=torch.uint8, device=like.device)
    dist = torch.empty(n, dtype=like.dtype, device=like.device)
    return simp, dist, f64


def _pairs(pairs, device, npoly=None, validate=False):
    """(N, 2) integer pair list -> two contiguous int32 index tensors.  With validate=True the
    indices are range-checked against the pool (one device reduction + a host read, i.e. a
    synchronisation; off by default on the zero-copy path -- out-of-range indices are undefined
    behaviour there exactly as in the reference's indexed kernel, gpu/opengjk_gpu.cu:1445-1631)."""
    torch = _torch()
    p = torch.as_tensor(pairs, device=device).to(torch.int32).reshape(-1, 2)
    if validate and npoly is not None and p.numel():
        lo, hi = int(p.min().item()), int(p.max().item())
        if lo < 0 or hi >= npoly:
            raise ValueError(f"pair index out of range: [{lo}, {hi}] vs {npoly} polytopes")
    return p[:, 0].contiguous(), p[:, 1].contiguous()


def compute_minimum_distance(bodies1, bodies2, counts1=None, counts2=None, engine=None):
    """Batched GJK over device tensors: pair p = (bodies1[p], bodies2[p]).
    bodies: (P, V, 3) CUDA tensors (same dtype); counts: optional (P,) true vertex counts."""
    torch = _torch()
    b1, b2 = _as_bodies(bodies1), _as_bodies(bodies2)
    if b1.shape[0] != b2.shape[0] or b1.dtype != b2.dtype or b1.device != b2.device:
        raise ValueError("bodies1/bodies2 must agree in batch size, dtype and device")
    eng = engine or api.default_engine(b1.device.index or 0)
    n = b1.shape[0]
    c1, o1, k1 = _layout(b1, counts1)
    c2, o2, k2 = _layout(b2, counts2)
    simp, dist, f64 = _outputs(n, b1)
    eng.gjk_device("f64" if f64 else "f32", n, c1, o1, k1, None, c2, o2, k2, None, simp, dist,
                   api.MODE_AUTO, torch.cuda.current_stream(b1.device).cuda_stream)
    return _soa(simp, dist, n, f64)


def compute_minimum_distance_indexed(polytopes, pairs, counts=None, engine=None, validate=False):
    """Indexed batch over one device pool (reference: compute_minimum_distance_indexed,
    gpu/include/openGJK_GPU.h:331-366): polytopes (P, V, 3), pairs (N, 2) integer tensor.
    validate=True range-checks the pair indices first (synchronises)."""
    torch = _torch()
    b = _as_bodies(polytopes)
    eng = engine or api.default_engine(b.device.index or 0)
    pa, pb = _pairs(pairs, b.device, b.shape[0], validate)
    n = int(pa.numel())
    c, o, k = _layout(b, counts)
    simp, dist, f64 = _outputs(n, b)
    if n:
        eng.gjk_device("f64" if f64 else "f32", n, c, o, k, pa, c, o, k, pb, simp, dist, api.MODE_AUTO,
                       torch.cuda.current_stream(b.device).cuda_stream)
    return _soa(simp, dist, n, f64)


class TensorStore:
    """A device pool kept across frames (ogjk_store: packed blocks + cluster tables for large
    bodies); `query(pairs)` runs GJK, `query(pairs, epa=True)` GJK + EPA, on torch's stream."""

    def __init__(self, polytopes, counts=None, engine=None):
        torch = _torch()
        b = _as_bodies(polytopes)
        self.engine = engine or api.default_engine(b.device.index or 0)
        self.device, self.dtype = b.device, b.dtype
        c, o, k = _layout(b, counts)
        self.store = self.engine.store_create(c, o, k, torch.cuda.current_stream(b.device).cuda_stream)

    def query(self, pairs, epa=False, validate=False):
        torch = _torch()
        pa, pb = _pairs(pairs, self.device, self.store.num_polytopes, validate)
        n = int(pa.numel())
        simp, dist, f64 = _outputs(n, torch.empty(0, dtype=self.dtype, device=self.device))
        st = torch.cuda.current_stream(self.device).cuda_stream
        if not epa:
            if n:
                self.engine.store_gjk(self.store, pa, self.store, pb, n, simp, dist, st)
            return _soa(simp, dist, n, f64)
        w1, w2, nr = (torch.empty((n, 3), dtype=self.dtype, device=self.device) for _ in range(3))
        if n:
            self.engine.store_gjk_epa(self.store, pa, self.store, pb, n, simp, dist, w1, w2, nr, st)
        out = _soa(simp, dist, n, f64)
        out.update({"witnesses1": w1, "witnesses2": w2, "contact_normals": nr, "penetration_depth": -dist.clamp(max=0)})
        return out

    # -- broad phase: boxes, candidate culling, all-pairs (ogjk_store_aabbs_* / ogjk_pairs_*) ----
    def aabbs(self):
        """(P, 6) tensor {min xyz, max xyz} of every body of the pool."""
        torch = _torch()
        boxes = torch.empty((self.store.num_polytopes, 6), dtype=self.dtype, device=self.device)
        self.engine.store_aabbs(self.store, boxes, torch.cuda.current_stream(self.device).cuda_stream)
        return boxes

    def _prec(self):
        return "f64" if self.dtype == _torch().float64 else "f32"

    def cull(self, pairs, margin=0.0, boxes=None):
        """Candidates (N, 2) -> the (M, 2) sub-list, order kept, whose boxes are within `margin` on every
        axis (a necessary condition for distance <= margin).  Reads back one integer (M)."""
        torch = _torch()
        boxes = self.aabbs() if boxes is None else boxes
        pa, pb = _pairs(pairs, self.device)
        n = int(pa.numel())
        out = torch.empty((2, max(n, 1)), dtype=torch.int32, device=self.device)
        count = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.engine.pairs_cull(self._prec(), boxes, boxes, n, pa, pb, margin, n, out[0], out[1], count,
                               torch.cuda.current_stream(self.device).cuda_stream)
        m = int(count.item())
        return out[:, :m].t().contiguous()

    def candidate_pairs(self, margin=0.0, capacity=None, boxes=None):
        """All (i < j) whose boxes are within `margin`, sorted by (i, j), as an (M, 2) int32 tensor."""
        torch = _torch()
        boxes = self.aabbs() if boxes is None else boxes
        P = self.store.num_polytopes
        st = torch.cuda.current_stream(self.device).cuda_stream
        count = torch.zeros(1, dtype=torch.int64, device=self.device)
        if capacity is None:  # counting pass first (the list is produced by a second call)
            self.engine.pairs_all(self._prec(), boxes, P, margin, 0, None, None, count, st)
            capacity = int(count.item())
        out = torch.empty((2, max(capacity, 1)), dtype=torch.int32, device=self.device)
        self.engine.pairs_all(self._prec(), boxes, P, margin, capacity, out[0], out[1], count, st)
        m = min(int(count.item()), capacity)
        return out[:, :m].t().contiguous()

    def close(self):
        self.store.close()
