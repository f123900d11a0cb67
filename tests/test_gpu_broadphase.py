"""GPU parity of the broad-phase-adjacent pair production (opengjk_b200/csrc/broadphase.cuh,
SURVEY section 8(f)-3) against a numpy restatement of its definition: boxes, the stable cull of
a candidate list and the sorted all-pairs list must match index for index; every pair the
cull drops must really be farther apart than the margin (checked with the GJK distances)."""
import numpy as np
import pytest

from opengjk_b200 import workloads as W, torch_api
from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu
PRECS = [("f64", np.float64), ("f32", np.float32)]


def _boxes_ref(w):
    out = np.empty((w.cnt.shape[0], 6), dtype=w.coords.dtype)
    for h in range(w.cnt.shape[0]):
        p = w.coords[w.off[h]:w.off[h] + w.cnt[h]]
        out[h, :3] = np.fmin.reduce(p, axis=0)
        out[h, 3:] = np.fmax.reduce(p, axis=0)
    return out


def _near_ref(b1, b2, margin):
    m = b1.dtype.type(margin)
    with np.errstate(invalid="ignore"):
        return (~((b1[:, :3] - b2[:, 3:]) > m)).all(1) & (~((b2[:, :3] - b1[:, 3:]) > m)).all(1)


def _store(w):
    """TensorStore over a ragged workload: bodies are padded to the longest one."""
    import torch
    P, V = w.cnt.shape[0], int(w.cnt.max())
    pad = np.zeros((P, V, 3), dtype=w.coords.dtype)
    for h in range(P):
        pad[h, :w.cnt[h]] = w.coords[w.off[h]:w.off[h] + w.cnt[h]]
    dev = torch.device("cuda:0")
    return torch_api.TensorStore(torch.from_numpy(pad).to(dev), torch.from_numpy(w.cnt).to(dev))


@pytest.mark.parametrize("prec,dt", PRECS)
def test_boxes_cull_and_all_pairs(prec, dt):
    import torch
    rng = np.random.default_rng(3)
    w = W.mixed_hulls(6000, seed=21, dtype=dt)          # 1..64-vertex bodies, ragged
    store = _store(w)
    boxes = store.aabbs()
    torch.cuda.synchronize()
    bref = _boxes_ref(w)
    assert boxes.cpu().numpy().tobytes() == bref.tobytes()

    P = w.cnt.shape[0]
    cand = rng.integers(0, P, size=(200_003, 2)).astype(np.int32)   # not a multiple of the block size
    for margin in (0.0, 0.25, 3.0):
        keep = _near_ref(bref[cand[:, 0]], bref[cand[:, 1]], margin)
        got = store.cull(torch.from_numpy(cand).cuda(), margin, boxes).cpu().numpy()
        assert np.array_equal(got, cand[keep]), margin
    # nothing that was dropped is within the margin: GJK distances of a sample of dropped pairs
    margin = 0.25
    keep = _near_ref(bref[cand[:, 0]], bref[cand[:, 1]], margin)
    dropped = cand[~keep][:5000]
    d, _ = Oracle(prec).batch(w.coords, w.off, w.cnt, dropped[:, 0].copy(), dropped[:, 1].copy())
    assert (d > margin).all()

    sub = 1500                                            # all-pairs on a sub-pool (host check is O(n^2))
    w2 = W.Workload("sub", w.coords[:w.off[sub]], w.off[:sub], w.cnt[:sub], w.pa[:1], w.pb[:1])
    s2 = _store(w2)
    i, j = np.triu_indices(sub, 1)
    for margin in (0.0, 0.5):
        near = _near_ref(bref[i], bref[j], margin)
        ref = np.stack([i[near], j[near]], 1).astype(np.int32)
        got = s2.candidate_pairs(margin).cpu().numpy()
        assert np.array_equal(got, ref), margin
        few = s2.candidate_pairs(margin, capacity=100).cpu().numpy()   # truncated output keeps the prefix
        assert np.array_equal(few, ref[:100])
    store.close(); s2.close()


def test_large_bodies_and_non_finite_boxes():
    import torch
    w = W.large_hull_pool(16, 40, 300, 3000, 9, dtype=np.float32)
    coords = w.coords.copy()
    coords[w.off[3] + 5, 1] = np.nan        # one NaN coordinate: skipped by the box (fmin/fmax)
    coords[w.off[7]:w.off[7] + w.cnt[7], 2] = np.nan   # a whole axis NaN: NaN box, never culled
    w = W.Workload("nf", coords, w.off, w.cnt, w.pa, w.pb)
    store = _store(w)
    boxes = store.aabbs()
    torch.cuda.synchronize()
    bref = _boxes_ref(w)
    assert np.array_equal(boxes.cpu().numpy(), bref, equal_nan=True)   # NaN payloads differ (CUDA canonical NaN)
    assert np.isnan(bref[7, [2, 5]]).all() and not np.isnan(bref[3]).any()
    P = w.cnt.shape[0]
    i, j = np.triu_indices(P, 1)
    near = _near_ref(bref[i], bref[j], 0.1)
    ref = np.stack([i[near], j[near]], 1).astype(np.int32)
    got = store.candidate_pairs(0.1).cpu().numpy()
    assert np.array_equal(got, ref)
    assert ((ref == 7).any(1).sum()) == P - 1            # the NaN box pairs with everybody
    # cull -> indexed GJK on the survivors, straight from device lists
    out = store.query(store.cull(torch.from_numpy(ref).cuda(), 0.1, boxes))
    torch.cuda.synchronize()
    assert out["distances"].shape[0] == ref.shape[0]
    store.close()


def test_empty_inputs():
    import torch
    w = W.mixed_hulls(10, seed=2, dtype=np.float32)
    store = _store(w)
    assert store.cull(torch.empty((0, 2), dtype=torch.int32, device="cuda:0")).shape == (0, 2)
    one = torch_api.TensorStore(torch.zeros((1, 4, 3), device="cuda:0"))
    assert one.candidate_pairs(1.0).shape == (0, 2)
    store.close(); one.close()
