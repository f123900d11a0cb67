"""GPU parity of the zero-copy tensor front door (opengjk_b200/torch_api.py, SURVEY section
8(f)-2): (P, V, 3) CUDA tensors / DLPack in, SoA tensors out, compared bit for bit with the
CPU oracle on the same ragged batches."""
import numpy as np
import pytest

from opengjk_b200 import api, torch_api
from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu
PRECS = [("f64", np.float64), ("f32", np.float32)]


def _ragged(P, V, seed, dt, scale=3.0):
    rng = np.random.default_rng(seed)
    pts = rng.normal(size=(P, V, 3))
    pts /= np.linalg.norm(pts, axis=2, keepdims=True)
    pts += scale * (rng.random((P, 1, 3)) - 0.5)
    cnt = rng.integers(4, V + 1, size=P).astype(np.int32)
    return pts.astype(dt), cnt


def _oracle(prec, A, ka, B, kb, pa, pb):
    P, V, _ = A.shape
    Q, W, _ = B.shape
    coords = np.concatenate([A.reshape(-1, 3), B.reshape(-1, 3)])
    off = np.concatenate([np.arange(P, dtype=np.int64) * V, P * V + np.arange(Q, dtype=np.int64) * W])
    cnt = np.concatenate([ka, kb]).astype(np.int32)
    return Oracle(prec).batch(coords, off, cnt, pa.astype(np.int32), (pb + P).astype(np.int32))


def _same(out, dref, sref):
    assert out["distances"].cpu().numpy().tobytes() == dref.tobytes()
    assert np.array_equal(out["simplex_nvrtx"].cpu().numpy(), sref["nvrtx"])
    assert out["simplex_vrtx_idx"].cpu().numpy().tobytes() == np.ascontiguousarray(sref["vrtx_idx"]).tobytes()
    assert out["simplex_vrtx"].cpu().numpy().tobytes() == np.ascontiguousarray(sref["vrtx"]).tobytes()
    assert out["witnesses1"].cpu().numpy().tobytes() == np.ascontiguousarray(sref["witnesses"][:, 0]).tobytes()
    assert out["witnesses2"].cpu().numpy().tobytes() == np.ascontiguousarray(sref["witnesses"][:, 1]).tobytes()
    eps = np.finfo(dref.dtype).eps
    assert np.array_equal(out["intersecting"].cpu().numpy(), ~(dref > eps))


@pytest.mark.parametrize("prec,dt", PRECS)
def test_pairwise_ragged_batch(prec, dt):
    import torch
    A, ka = _ragged(3000, 40, 1, dt)
    B, kb = _ragged(3000, 24, 2, dt)
    ar = np.arange(3000)
    dref, sref = _oracle(prec, A, ka, B, kb, ar, ar)
    dev = torch.device("cuda:0")
    out = torch_api.compute_minimum_distance(torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev),
                                             torch.from_numpy(ka).to(dev), torch.from_numpy(kb).to(dev))
    torch.cuda.synchronize()
    _same(out, dref, sref)
    # full rows (no counts) through DLPack exporters
    dref, sref = _oracle(prec, A, np.full(3000, 40), B, np.full(3000, 24), ar, ar)

    class Exporter:  # any object with __dlpack__ / __dlpack_device__
        def __init__(self, t):
            self.t = t

        def __dlpack__(self, stream=None):
            return self.t.__dlpack__(stream=stream)

        def __dlpack_device__(self):
            return self.t.__dlpack_device__()

    out = torch_api.compute_minimum_distance(Exporter(torch.from_numpy(A).to(dev)), Exporter(torch.from_numpy(B).to(dev)))
    torch.cuda.synchronize()
    _same(out, dref, sref)


@pytest.mark.parametrize("prec,dt", PRECS)
def test_indexed_and_store(prec, dt):
    import torch
    rng = np.random.default_rng(5)
    A, ka = _ragged(500, 700, 3, dt, scale=2.0)   # some bodies large enough for cluster tables in the store
    pairs = rng.integers(0, 500, size=(4000, 2))
    dref, sref = _oracle(prec, A, ka, A[:0], ka[:0], pairs[:, 0], pairs[:, 1] - 500)
    dev = torch.device("cuda:0")
    tA, tk, tp = torch.from_numpy(A).to(dev), torch.from_numpy(ka).to(dev), torch.from_numpy(pairs).to(dev)
    out = torch_api.compute_minimum_distance_indexed(tA, tp, tk)
    torch.cuda.synchronize()
    _same(out, dref, sref)
    store = torch_api.TensorStore(tA, tk)
    assert store.store.table_bytes > 0
    out = store.query(tp)
    torch.cuda.synchronize()
    _same(out, dref, sref)
    # GJK + EPA on the same store: separated pairs keep the GJK answer, penetrating ones get a depth
    e = store.query(tp, epa=True)
    torch.cuda.synchronize()
    d = e["distances"].cpu().numpy()
    sep = dref > np.finfo(dt).eps
    assert d[sep].tobytes() == dref[sep].tobytes()
    assert (d[~sep] <= 0).all() and np.isfinite(e["contact_normals"].cpu().numpy()).all()
    ref = api.default_engine().gjk_epa_host(A.reshape(-1, 3), np.arange(500, dtype=np.int64) * 700, ka,
                                            pairs[:, 0].astype(np.int32), pairs[:, 1].astype(np.int32))
    assert d.tobytes() == ref["distances"].tobytes()
    assert e["witnesses1"].cpu().numpy().tobytes() == ref["witness1"].tobytes()
    store.close()


def test_intersect_flags_entry():
    """ogjk_intersect_flags_*: the device-side flag equals the rule applied to the oracle's distances."""
    import torch
    for prec, dt in PRECS:
        A, ka = _ragged(4000, 20, 11, dt, scale=2.5)
        B, kb = _ragged(4000, 20, 12, dt, scale=2.5)
        ar = np.arange(4000)
        dref, _ = _oracle(prec, A, ka, B, kb, ar, ar)
        dev = torch.device("cuda:0")
        out = torch_api.compute_minimum_distance(torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev),
                                                 torch.from_numpy(ka).to(dev), torch.from_numpy(kb).to(dev))
        flags = torch.zeros(4000, dtype=torch.uint8, device=dev)
        api.default_engine().intersect_flags(prec, 4000, out["distances"], flags, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        want = ~(dref > np.finfo(dt).eps)
        assert 0.05 < want.mean() < 0.95
        assert np.array_equal(flags.cpu().numpy().astype(bool), want)
        assert np.array_equal(out["intersecting"].cpu().numpy(), want)


def test_async_on_side_stream():
    """Calls are enqueued on torch's current stream; no hidden synchronisation is needed for correctness."""
    import torch
    A, ka = _ragged(2000, 16, 7, np.float32)
    B, kb = _ragged(2000, 16, 8, np.float32)
    ar = np.arange(2000)
    dref, sref = _oracle("f32", A, ka, B, kb, ar, ar)
    dev = torch.device("cuda:0")
    s = torch.cuda.Stream(dev)
    with torch.cuda.stream(s):
        tA, tB = torch.from_numpy(A).to(dev, non_blocking=True), torch.from_numpy(B).to(dev, non_blocking=True)
        out = torch_api.compute_minimum_distance(tA, tB, torch.from_numpy(ka).to(dev), torch.from_numpy(kb).to(dev))
        host = {k: v.cpu() for k, v in out.items()}
    s.synchronize()
    assert host["distances"].numpy().tobytes() == dref.tobytes()
    assert np.array_equal(host["simplex_nvrtx"].numpy(), sref["nvrtx"])
