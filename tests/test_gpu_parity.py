"""GPU parity: the CUDA kernels, called through the C ABI, against the CPU
oracle and the committed outputs of the unmodified reference.

Bar (BASELINE.json north_star): distance/witness within 1e-10 (fp64) / 1e-5
(fp32) relative, flags and support indices bit-exact.  Because the kernels keep
the reference's operation order un-fused, the tests assert the stronger
property: every byte of distance and gkSimplex equal."""
import os

import numpy as np
import pytest

from opengjk_b200 import api, workloads as W
from oracle.pyoracle import Oracle, Reference
from tests.golden.make_golden import checksum, golden_workloads
from tests.test_oracle import ANALYTIC, rotated_cubes

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
PRECS = [("f64", np.float64), ("f32", np.float32)]
REL = {"f64": 1e-10, "f32": 1e-5}


@pytest.fixture(scope="module")
def eng():
    e = api.Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def eng_exp():
    """Engine over libopengjk_b200_exp.so: same ABI plus the measured-and-lost kernel flavours."""
    from opengjk_b200 import _lib
    e = api.Engine(0, lib_path=_lib.EXP_LIB_PATH)
    yield e
    e.close()


def _assert_same(name, d, s, dref, sref, prec):
    # tolerance form of the bar first (readable failure), then byte equality
    ok = ~np.isnan(dref)
    np.testing.assert_allclose(d[ok], dref[ok], rtol=REL[prec], atol=0, err_msg=name)
    assert np.array_equal(s["nvrtx"], sref["nvrtx"]), name
    assert np.array_equal((d > np.finfo(d.dtype).eps), (dref > np.finfo(d.dtype).eps)), name  # EPA gate flag
    assert d.tobytes() == dref.tobytes(), f"{name}: distance bits differ"
    # field-wise byte equality (struct padding is not part of the contract)
    bad = np.zeros(len(s), dtype=bool)
    for f in ("nvrtx", "vrtx", "vrtx_idx", "witnesses"):
        a = np.ascontiguousarray(s[f]).view(np.uint8).reshape(len(s), -1)
        b = np.ascontiguousarray(sref[f]).view(np.uint8).reshape(len(s), -1)
        bad |= (a != b).any(1)
    bad = np.nonzero(bad)[0]
    assert bad.size == 0, f"{name}: {bad.size} simplices differ, first {bad[:5]}"


def _gpu(eng, w, mode=api.MODE_AUTO, threshold=None):
    """Device-resident call through ogjk_gjk_device_* (torch only holds the memory)."""
    import torch
    prec = "f64" if w.coords.dtype == np.float64 else "f32"
    sdt = api.SIMPLEX_F64 if prec == "f64" else api.SIMPLEX_F32
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    coords, off, cnt, pa, pb = t(w.coords.reshape(-1)), t(w.off), t(w.cnt), t(w.pa), t(w.pb)
    simp = torch.zeros(w.npairs * sdt.itemsize, dtype=torch.uint8, device=dev)
    dist = torch.zeros(w.npairs, dtype=coords.dtype, device=dev)
    if threshold:
        eng.set_warp_threshold(threshold)
    eng.gjk_device(prec, w.npairs, coords, off, cnt, pa, coords, off, cnt, pb, simp, dist, mode,
                   torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    if threshold:
        eng.set_warp_threshold(192)
    return dist.cpu().numpy(), simp.cpu().numpy().view(sdt).reshape(-1)


@pytest.mark.parametrize("prec,dt", PRECS)
def test_committed_reference_outputs(eng, prec, dt):
    gold = np.load(os.path.join(GOLD, f"gjk_golden_{prec}.npz"))
    for w in golden_workloads(dt):
        assert checksum(w) == bytes(gold[w.name + "/sha"]).decode()
        sref = gold[w.name + "/simplex"].copy().view(api.SIMPLEX_F64 if prec == "f64" else api.SIMPLEX_F32).reshape(-1)
        for mode in (api.MODE_AUTO, api.MODE_THREAD, api.MODE_WARP):
            d, s = _gpu(eng, w, mode)
            _assert_same(f"{w.name}/mode{mode}", d, s, gold[w.name + "/dist"], sref, prec)


@pytest.mark.parametrize("prec,dt", PRECS)
def test_config1_example(eng, prec, dt):
    d, s = api.single_query(W.USER_P, W.USER_Q, dt, eng)
    assert f"{d:f}" == "3.653650"  # README.md:55-59
    assert int(s["nvrtx"]) == 2 and s["vrtx_idx"][:2].tolist() == [[6, 6], [1, 1]]
    if prec == "f64":
        assert d == 3.6536497222945008
    else:
        assert np.float32(d) == np.float32(3.6536495685577393)


@pytest.mark.parametrize("prec,dt", PRECS)
@pytest.mark.parametrize("maker", [
    lambda dt: W.mixed_hulls(200000, seed=101, dtype=dt),
    lambda dt: W.boxes(200000, True, 102, dtype=dt),
    lambda dt: W.boxes(200000, False, 103, dtype=dt),
    lambda dt: W.random_hulls(100000, 8, 64, 104, 1.0, dt, 0.5),
    lambda dt: W.tiny_bodies(400, 105, 1.0, dt),
    lambda dt: W.tiny_bodies(400, 106, 1e4, dt),
    lambda dt: W.large_hull_pool(2000, 64, 1000, 4000, 107, dtype=dt),
    lambda dt: W.random_hulls(20000, 1, 300, 108, 3.0, dt),
], ids=["mixed", "boxes_oriented", "boxes_aligned", "intersecting", "tiny", "tiny_1e4", "large_pool", "ragged_1_300"])
def test_random_batches_vs_oracle(eng, prec, dt, maker):
    w = maker(dt)
    o = Oracle(prec)
    dref, sref = o.batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=os.cpu_count() or 1)
    d, s = _gpu(eng, w)
    _assert_same(w.name, d, s, dref, sref, prec)


DEFAULT_CLASSES = [(4, 0), (32, 2), (128, 8), (2048, 32)]


@pytest.mark.parametrize("prec,dt", PRECS)
def test_scheduler_classes_do_not_change_results(eng, eng_exp, prec, dt):
    """Any lanes-per-pair choice (1..32, register-resident or not) gives the same bytes; the
    experimental flavours (staged, persistent, scan-assist, CTA-per-pair) are swept in the
    -DOGJK_EXPERIMENTAL build and rejected by the product library."""
    with pytest.raises(api.OpenGJKError):
        eng.set_classes([(80, 108), (2048, 32)])
    with pytest.raises(api.OpenGJKError):
        eng.set_classes([(4, 0)])  # the last class is open-ended: it cannot be the register-resident kernel
    plain = eng
    eng = eng_exp
    w = W.random_hulls(30000, 1, 150, 55, 3.0, dt)
    base = _gpu(eng, w, api.MODE_THREAD)
    tables = [[(2048, g)] for g in (1, 2, 4, 8, 16, 32)] + [
        [(4, 0), (2048, 1)], [(4, 0), (8, 2), (16, 4), (32, 8), (64, 16), (2048, 32)], [(3, 1), (5, 32), (2048, 2)]]
    # shared-memory staged flavours (100 + lanes), one and several slot sizes
    tables += [[(80, 100 + g), (2048, g)] for g in (1, 2, 4, 8, 16, 32)]
    tables += [[(4, 0), (12, 102), (24, 104), (48, 108), (80, 132), (2048, 16)]]
    # persistent-lane flavours (200 + lanes; 200 = register-resident)
    tables += [[(2048, 200 + g)] for g in (1, 2, 4, 8, 16, 32)]
    tables += [[(4, 200), (24, 202), (64, 204), (2048, 232)]]
    # scan-assist flavours (300 + lanes of state per pair; scans use the whole warp)
    tables += [[(2048, 300 + g)] for g in (2, 4, 8, 16)]
    # CTA-per-pair flavours (500 + warps), shared memory sized by the class bound
    tables += [[(80, 500 + g), (2048, 32)] for g in (2, 4, 8, 16)]
    try:
        for t in tables:
            eng.set_classes(t)
            d, s = _gpu(eng, w, api.MODE_AUTO)
            _assert_same(f"classes{t}", d, s, base[0], base[1], prec)
        for t in tables[:9]:  # the plain flavours through the product library as well
            plain.set_classes(t)
            d, s = _gpu(plain, w, api.MODE_AUTO)
            _assert_same(f"product classes{t}", d, s, base[0], base[1], prec)
    finally:
        eng.set_classes(DEFAULT_CLASSES)
        plain.set_classes(DEFAULT_CLASSES)


@pytest.mark.parametrize("prec,dt", PRECS)
def test_packed_store_api(eng, prec, dt):
    """ogjk_store_create (host and device sources) + ogjk_store_gjk, plain and indexed."""
    import torch
    o = Oracle(prec)
    sdt = api.SIMPLEX_F64 if prec == "f64" else api.SIMPLEX_F32
    dev = torch.device("cuda:0")
    for w in (W.mixed_hulls(20000, seed=31, dtype=dt), W.boxes(20000, True, 32, dtype=dt),
              W.large_hull_pool(3000, 64, 300, 1200, 33, dtype=dt), W.tiny_bodies(100, 34, 1.0, dt)):
        dref, sref = o.batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=os.cpu_count() or 1)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        for source in ("host", "device"):
            if source == "host":
                store = eng.store_create(w.coords, w.off, w.cnt)
            else:
                store = eng.store_create(t(w.coords.reshape(-1)), t(w.off), t(w.cnt),
                                         torch.cuda.current_stream().cuda_stream)
            assert store.num_polytopes == w.npoly and store.nbytes >= w.coords.nbytes
            simp = torch.zeros(w.npairs * sdt.itemsize, dtype=torch.uint8, device=dev)
            dist = torch.zeros(w.npairs, dtype=t(w.coords[:1]).dtype, device=dev)
            eng.store_gjk(store, t(w.pa), store, t(w.pb), w.npairs, simp, dist, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            _assert_same(f"{w.name}/{source}", dist.cpu().numpy(), simp.cpu().numpy().view(sdt).reshape(-1), dref, sref,
                         prec)
            store.close()


def test_live_reference_library_if_present(eng):
    if not Reference.available("f64"):
        pytest.skip("oracle/_ref not shipped")
    for prec, dt in PRECS:
        w = W.mixed_hulls(50000, seed=77, dtype=dt)
        dref, sref = Reference(prec).batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=os.cpu_count() or 1)
        d, s = _gpu(eng, w)
        _assert_same("live_ref", d, s, dref, sref, prec)


@pytest.mark.parametrize("case", ANALYTIC, ids=lambda c: c[0])
def test_analytic_distances_fp64(eng, case):
    _, a, b, exp, _ = case
    d, _ = api.single_query(a, b, np.float64, eng)
    assert np.isclose(d, exp, atol=1e-12)


def test_rotated_cubes_gap_sweep(eng):
    cubes = rotated_cubes()
    A, B, want = [], [], []
    for c0 in range(4):
        for c1 in range(4):
            dx = cubes[c0][:, 0].max() - cubes[c1][:, 0].min()
            for delta in (1e8, 1.0, 1e-4, 1e-8, 1e-12):
                A.append(cubes[c0]); B.append(cubes[c1] + np.array([dx + delta, 0, 0])); want.append(delta)
    r = api.compute_minimum_distance(np.stack(A), np.stack(B), np.float64, eng)
    assert np.allclose(r["distances"], want)


@pytest.mark.parametrize("prec,dt", PRECS)
def test_host_entry_points_agree(eng, prec, dt):
    """Flat host API, struct-array API (plain + indexed) and the reference-style
    dict API all give the oracle's bytes."""
    w = W.mixed_hulls(3000, seed=9, dtype=dt)
    o = Oracle(prec)
    dref, sref = o.batch(w.coords, w.off, w.cnt, w.pa, w.pb)
    d, s = eng.run_workload(w)
    _assert_same("flat_host", d, s, dref, sref, prec)
    A = [w.polytope(h) for h in w.pa]
    B = [w.polytope(h) for h in w.pb]
    d, s = api.compute_minimum_distance_structs(A, B, dt, eng)
    _assert_same("struct_arrays", d, s, dref, sref, prec)
    bodies = [w.polytope(h) for h in range(w.npoly)]
    d, s = api.compute_minimum_distance_indexed_structs(bodies, np.stack([w.pa, w.pb], 1), dt, eng)
    _assert_same("struct_indexed", d, s, dref, sref, prec)
    r = api.compute_minimum_distance(A, B, dt, eng)
    assert r["distances"].tobytes() == dref.tobytes()
    assert np.array_equal(r["simplex_nvrtx"], sref["nvrtx"])
    r = api.compute_minimum_distance_indexed(bodies, np.stack([w.pa, w.pb], 1), dt, eng)
    assert r["distances"].tobytes() == dref.tobytes()


def test_empty_and_degenerate_batches(eng):
    # n == 0 is a silent no-op like the reference (gpu/opengjk_gpu.cu:2984)
    r = api.compute_minimum_distance(np.zeros((0, 4, 3), np.float32), np.zeros((0, 4, 3), np.float32), engine=eng)
    assert r["distances"].shape == (0,)
    # identical single points: distance 0
    d, s = api.single_query([[1, 2, 3]], [[1, 2, 3]], np.float64, eng)
    assert d == 0.0 and int(s["nvrtx"]) == 1
    # one big hull vs a point (ragged extremes in one batch)
    rng = np.random.default_rng(0)
    big = rng.normal(size=(5000, 3)); big /= np.linalg.norm(big, axis=1, keepdims=True)
    o = Oracle("f64")
    dref, sref, _ = o.gjk(big, [[3.0, 0, 0]])
    d, s = api.single_query(big, [[3.0, 0, 0]], np.float64, eng)
    assert d == dref and s.tobytes() == sref.tobytes()


def test_full_size_config2_properties(eng):
    """BASELINE config #2 at full size (1 M pairs, fp64): size-independent
    properties -- witness gap equals distance, distances of separated/overlapping
    slices have the right sign structure -- plus oracle equality on a strided sample."""
    w = W.mixed_hulls(1_000_000, seed=42, dtype=np.float64)
    d, s = _gpu(eng, w)
    gap = np.linalg.norm(s["witnesses"][:, 0] - s["witnesses"][:, 1], axis=1)
    ok = ~np.isnan(d)
    assert ok.mean() > 0.999
    assert np.all(np.abs(gap[ok] - d[ok]) <= 1e-6)  # reference check_witnesses uses 1e-5
    assert (d[900_000:] > 0).mean() > 0.9 and d[900_000:].mean() > d[:800_000].mean()  # centre scale 10: mostly separated
    assert (s["nvrtx"][800_000:900_000] == 4).mean() > 0.95   # centre scale 1: intersecting
    sel = np.arange(0, w.npairs, 37)
    o = Oracle("f64")
    dref, sref = o.batch(w.coords, w.off, w.cnt, w.pa[sel], w.pb[sel], nthreads=os.cpu_count() or 1)
    _assert_same("config2_sample", d[sel], s[sel], dref, sref, "f64")


@pytest.mark.parametrize("prec,dt", PRECS)
def test_chunk_pipelined_host_path(eng, prec, dt):
    """Un-indexed two-pool form (the reference's bd1[] / bd2[] arrays): large batches are cut into
    chunks whose H2D / kernels / D2H overlap; results must not depend on the chunking."""
    w = W.mixed_hulls(100_000, seed=61, dtype=dt)
    dref, sref = Oracle(prec).batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=os.cpu_count() or 1)
    pools = w.two_pools()
    try:
        for chunks in (8, 3, 16, 0):
            eng.set_pipeline_chunks(chunks)
            d, s = eng.gjk_host_pools(*pools)
            _assert_same(f"chunks{chunks}", d, s, dref, sref, prec)
    finally:
        eng.set_pipeline_chunks(8)
    # GJK + EPA through the same pipeline
    w = W.random_hulls(60_000, 8, 64, 62, 1.5, dt, 0.5)
    ref = Oracle(prec).gjk_epa_batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=os.cpu_count() or 1)
    got = eng.gjk_host_pools(*w.two_pools(), epa=True)
    for k in ("distances", "witness1", "witness2", "normals"):
        assert np.ascontiguousarray(got[k]).tobytes() == np.ascontiguousarray(ref[k]).tobytes(), k


@pytest.mark.parametrize("prec,dt", PRECS)
def test_multi_pass_scheduling_is_invisible(eng_exp, prec, dt):
    """Pairs parked after a few iterations and resumed in later passes give the one-pass bytes, for
    every split of the iteration budget (including parking after every single iteration).
    Experimental build only (the flavour lost to the one-pass kernel on B200)."""
    eng = eng_exp
    o = Oracle(prec)
    ws = [W.mixed_hulls(30000, seed=71, dtype=dt), W.random_hulls(6000, 40, 200, 72, 3.0, dt), W.tiny_bodies(60, 73, 1.0, dt)]
    refs = [o.batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=os.cpu_count() or 1) for w in ws]
    try:
        for passes, k1, k2 in ((2, 4, 3), (3, 4, 3), (2, 1, 1), (3, 1, 1), (3, 2, 20), (2, 23, 1)):
            eng.set_passes(passes, k1, k2)
            for w, (dref, sref) in zip(ws, refs):
                d, s = _gpu(eng, w)
                _assert_same(f"{w.name}/passes{passes}:{k1}:{k2}", d, s, dref, sref, prec)
    finally:
        eng.set_passes(1)


@pytest.mark.parametrize("prec,dt", PRECS)
def test_in_cta_regrouping_is_invisible(eng, eng_exp, prec, dt):
    """Pairs packed into other lanes of their CTA between iterations (regroup_kernel.cuh) give the plain
    group kernel's bytes for every choice of regroup points, including batches that do not fill a CTA.
    Experimental build only (measured slower on B200); the product library rejects the knob."""
    o = Oracle(prec)
    from tests.test_oracle import _cap_workload
    ws = [W.mixed_hulls(30000, seed=81, dtype=dt), W.mixed_hulls(77, seed=82, dtype=dt), W.tiny_bodies(60, 83, 1.0, dt),
          _cap_workload(dt, 300, 84)]
    refs = [o.batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=os.cpu_count() or 1) for w in ws]
    try:
        eng_exp.set_classes([(2048, 2)])  # every pair through the 2-lane class
        for r1, r2 in ((4, 6), (3, 5), (1, 2), (4, 0), (24, 25), (2, 20)):
            eng_exp.set_regroup(r1, r2)
            for w, (dref, sref) in zip(ws, refs):
                d, s = _gpu(eng_exp, w)
                _assert_same(f"{w.name}/regroup{r1}:{r2}", d, s, dref, sref, prec)
    finally:
        eng_exp.set_regroup(0)
        eng_exp.set_classes(DEFAULT_CLASSES)
    with pytest.raises(Exception):
        eng.set_regroup(4, 6)
    eng.set_regroup(0)


@pytest.mark.parametrize("prec,dt", PRECS)
def test_bulk_copy_staging_is_invisible(eng_exp, prec, dt):
    """The TMA-staged flavour (tma_kernel.cuh: cp.async.bulk + mbarrier, double buffered per warp) gives the
    plain group kernel's bytes -- for buffers that hold whole batches, buffers that hold only some of a
    batch's pairs (the rest is read from global memory) and batches with a ragged tail."""
    o = Oracle(prec)
    ws = [W.mixed_hulls(30000, seed=91, dtype=dt), W.mixed_hulls(77, seed=92, dtype=dt), W.tiny_bodies(60, 93, 1.0, dt),
          W.random_hulls(3000, 40, 300, 94, 3.0, dt)]
    refs = [o.batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=os.cpu_count() or 1) for w in ws]
    saved = os.environ.get("OGJK_TMA_SLOT")
    try:
        for slot in ("14336", "2048", "26624"):
            os.environ["OGJK_TMA_SLOT"] = slot
            for lanes in (602, 604, 608):
                eng_exp.set_classes([(2048, lanes)])
                for w, (dref, sref) in zip(ws, refs):
                    d, s = _gpu(eng_exp, w)
                    _assert_same(f"{w.name}/tma{lanes}/slot{slot}", d, s, dref, sref, prec)
    finally:
        if saved is None:
            os.environ.pop("OGJK_TMA_SLOT", None)
        else:
            os.environ["OGJK_TMA_SLOT"] = saved
        eng_exp.set_classes(DEFAULT_CLASSES)


def test_chunked_path_falls_back_on_scattered_layouts(eng):
    """Pools whose polytopes are NOT stored in pair order: the chunked pipeline's device-side layout
    check must send the batch to the one-shot path -- both when whole chunks overlap (caught on the
    host) and when only a few polytopes inside chunks are out of place (caught only on the device)."""
    dt = np.float64
    w = W.mixed_hulls(50_000, seed=63, dtype=dt)
    dref, sref = Oracle("f64").batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=os.cpu_count() or 1)
    c1, o1, k1, c2, o2, k2 = w.two_pools()
    rng = np.random.default_rng(64)

    def relocate(c, o, k, movers):
        """Same polytopes, same pair order, but the vertices of `movers` are stored at the end of the pool."""
        c = c.copy(); o = o.copy()
        tail = [c]
        end = c.shape[0]
        for h in movers:
            tail.append(c[o[h]:o[h] + k[h]].copy())
            c[o[h]:o[h] + k[h]] = 1e30            # poison the old place: reading it would change the result
            o[h] = end
            end += k[h]
        return np.concatenate(tail), o

    few = rng.choice(len(k1), 7, replace=False)                      # a handful inside chunks
    many = rng.permutation(len(k1))[: len(k1) // 2]                  # half of the pool
    for movers in (few, many):
        cc1, oo1 = relocate(c1, o1, k1, movers)
        cc2, oo2 = relocate(c2, o2, k2, movers[::-1])
        d, s = eng.gjk_host_pools(cc1, oo1, k1, cc2, oo2, k2)
        _assert_same(f"scattered{len(movers)}", d, s, dref, sref, "f64")


@pytest.mark.parametrize("prec,dt", PRECS)
def test_scheduler_edge_cases(eng, prec, dt):
    """Batch sizes around warp/block boundaries, one pair, a hull beyond the largest sort key
    (> 8192 vertices per pair), and a batch that populates every size class at once."""
    o = Oracle(prec)
    rng = np.random.default_rng(91)
    for n in (1, 31, 32, 33, 127, 129, 2049):
        w = W.random_hulls(n, 1, 40, 92 + n, 3.0, dt)
        dref, sref = o.batch(w.coords, w.off, w.cnt, w.pa, w.pb)
        d, s = _gpu(eng, w)
        _assert_same(f"n{n}", d, s, dref, sref, prec)
    bodies_a, bodies_b = [], []
    for n1, n2 in ((20000, 3), (5, 9000), (8, 8), (4, 4), (1, 1), (70, 60), (300, 200), (1000, 1200), (3, 700)):
        a = rng.normal(size=(n1, 3)); a /= np.linalg.norm(a, axis=1, keepdims=True)
        b = rng.normal(size=(n2, 3)); b /= np.linalg.norm(b, axis=1, keepdims=True)
        bodies_a.append(a); bodies_b.append(b + rng.uniform(-2.5, 2.5, size=3))
    w = W.from_bodies(bodies_a * 40, bodies_b * 40, dt, "all_classes")
    dref, sref = o.batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=os.cpu_count() or 1)
    d, s = _gpu(eng, w)
    _assert_same("all_classes", d, s, dref, sref, prec)
    d, s = eng.run_workload(w)
    _assert_same("all_classes_host", d, s, dref, sref, prec)


@pytest.mark.parametrize("prec,dt", PRECS)
def test_iteration_cap_exit_every_kernel_flavour(eng, eng_exp, prec, dt):
    """>= 1000 pairs that end on the iteration cap (scalar/openGJK.c:1036-1042, k == mk:
    witnesses of a non-converged simplex) through every kernel flavour, byte-equal to the
    oracle and -- when oracle/_ref travelled -- to the unmodified reference library."""
    from tests.test_oracle import _cap_workload
    w = _cap_workload(dt, 6000, 12)
    dref, sref, it = Oracle(prec).batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=os.cpu_count() or 1, want_iters=True)
    assert int((it >= 25).sum()) >= 1000
    if Reference.available(prec):
        saved = os.dup(1)
        null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(null, 1)
        try:
            dr, sr = Reference(prec).batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=os.cpu_count() or 1)
        finally:
            os.dup2(saved, 1)
            os.close(saved)
            os.close(null)
        _assert_same("cap/oracle-vs-reference", dref, sref, dr, sr, prec)
    for mode in (api.MODE_AUTO, api.MODE_THREAD, api.MODE_WARP):
        d, s = _gpu(eng, w, mode)
        _assert_same(f"cap/mode{mode}", d, s, dref, sref, prec)
    tables = [[(2048, g)] for g in (1, 2, 4, 8, 16, 32)]
    exp_tables = [[(2048, 100 + 4)], [(2048, 200 + 2)], [(2048, 300 + 8)], [(2048, 500 + 4)]]
    try:
        for t in tables:
            eng.set_classes(t)
            d, s = _gpu(eng, w)
            _assert_same(f"cap/classes{t}", d, s, dref, sref, prec)
        for t in tables[:3] + exp_tables:
            eng_exp.set_classes(t)
            d, s = _gpu(eng_exp, w)
            _assert_same(f"cap/exp classes{t}", d, s, dref, sref, prec)
        eng_exp.set_classes([(2048, 2)])
        eng_exp.set_passes(3, 4, 3)
        d, s = _gpu(eng_exp, w)
        _assert_same("cap/passes", d, s, dref, sref, prec)
    finally:
        eng.set_classes(DEFAULT_CLASSES)
        eng_exp.set_classes(DEFAULT_CLASSES)
        eng_exp.set_passes(1)
    # host entry points (chunked and one-shot pipelines) and the packed store
    d, s = eng.gjk_host_pools(*w.two_pools())
    _assert_same("cap/host pools", d, s, dref, sref, prec)


@pytest.mark.parametrize("prec,dt", PRECS)
def test_struct_array_entry_points_with_pipelined_gather(eng, prec, dt):
    """Arrays of gkPolytope large enough for the host-thread gather (csrc/abi.cu: stage_structs,
    HostPool) to run chunk-pipelined against the uploads: plain, indexed and GJK+EPA forms,
    repeated calls on one context, all byte-equal to the oracle."""
    w = W.mixed_hulls(40000, seed=81, dtype=dt)
    o = Oracle(prec)
    dref, sref = o.batch(w.coords, w.off, w.cnt, w.pa, w.pb, nthreads=os.cpu_count() or 1)
    A = [w.polytope(h) for h in w.pa]
    B = [w.polytope(h) for h in w.pb]
    for _ in range(2):
        d, s = api.compute_minimum_distance_structs(A, B, dt, eng)
        _assert_same("structs", d, s, dref, sref, prec)
    pool = [w.polytope(h) for h in range(w.npoly)]
    d, s = api.compute_minimum_distance_indexed_structs(pool, np.stack([w.pa, w.pb], 1), dt, eng)
    _assert_same("structs/indexed", d, s, dref, sref, prec)
    we = W.random_hulls(20000, 8, 64, 82, 1.0, dt, 0.5)
    ref = o.gjk_epa_batch(we.coords, we.off, we.cnt, we.pa, we.pb, nthreads=os.cpu_count() or 1)
    d, s, w1, w2 = api.compute_gjk_epa_structs([we.polytope(h) for h in we.pa], [we.polytope(h) for h in we.pb], dt, eng)
    assert d.tobytes() == ref["distances"].tobytes()
    assert w1.tobytes() == ref["witness1"].tobytes() and w2.tobytes() == ref["witness2"].tobytes()
