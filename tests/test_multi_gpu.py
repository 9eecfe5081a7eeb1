"""The row-sharded search behind the C ABI (vdb_init + one vdb_flat_knn call; csrc/multi.cu).

Runs on ONE GPU: a device may be registered several times, so three uneven shards on device 0 exercise every phase of
the sharded call (slice upload + broadcast, global thresholds from the gathered sample scores, filter + rerank,
owner-sliced merge over "peer" memory, completeness check, exact re-scan of flagged queries). With two or more GPUs
visible the same cases also run with one shard per GPU.

Parity: the sharded result must be BIT-identical to the unsharded exact scan (ids, distance bits, counts) and must
satisfy the oracle parity rule (reference FlatIndex::knn, src/index_algorithm/flat_index.rs:48-57).
"""
import ctypes as C

import numpy as np
import pytest

from conftest import assert_knn_parity

pytestmark = pytest.mark.gpu


def _data(n, nq, dim, seed=0, dtype=np.float32):
    rng = np.random.default_rng(seed)
    proto = rng.random((256, dim), dtype=np.float32) * 0.15
    base = (proto[rng.integers(0, 256, n)] + 0.02 * rng.standard_normal((n, dim), dtype=np.float32)).clip(0, 1)
    q = (proto[rng.integers(0, 256, nq)] + 0.02 * rng.standard_normal((nq, dim), dtype=np.float32)).clip(0, 1)
    if dtype == np.uint8:
        return (base * 255).astype(np.uint8), (q * 255).astype(np.uint8)
    return base.astype(np.float32), q.astype(np.float32)


def _device_sets():
    import torch
    sets = [[0, 0, 0]]
    if torch.cuda.device_count() >= 2:
        sets.append(list(range(min(torch.cuda.device_count(), 4))))
    return sets


@pytest.fixture()
def vdb():
    import lab_1806_vec_db_b200 as V
    yield V
    V.init_devices([])
    V._lib.check(V.lib().vdb_debug_force_redo(0))
    V._lib.check(V.lib().vdb_flat_set_path(0))


def _same(got, want):
    assert (got[0].astype(np.int64) == want[0].astype(np.int64)).all()
    assert (np.asarray(got[1]).view(np.uint32) == np.asarray(want[1]).view(np.uint32)).all()
    assert (np.asarray(got[2]) == np.asarray(want[2])).all()


@pytest.mark.parametrize("metric,dtype", [("l2sqr", np.float32), ("cosine", np.float32), ("l2sqr", np.uint8)])
def test_one_call_on_a_sharded_set_equals_the_unsharded_scan(vdb, oracle, metric, dtype):
    V = vdb
    n, nq, dim = 300_000, 300, 128
    base, q = _data(n, nq, dim, 1, dtype)
    base[77_777] = base[123]            # an exact duplicate in another shard: the lower id must come first
    V.init_devices([])
    full = V.FlatIndex.from_vec_set(base, metric)
    full.vec_set.set_flat_path("scan")
    want = {k: full.knn_batch(q, k) for k in (10, 100)}
    for devs in _device_sets():
        V.init_devices(devs)
        idx = V.FlatIndex.from_vec_set(base, metric)          # vdb_dataset_create: row-sharded over `devs`
        assert [d for d, _, _ in idx.vec_set.shards()] == devs
        assert idx.vec_set.shards()[-1][2] == n and len(idx.vec_set) == n
        for k in (10, 100):
            _same(idx.knn_batch(q, k), want[k])                # tensor phases, global thresholds
            _same(idx.knn_batch(q[:5], k), tuple(a[:5] for a in want[k]))    # small batch: exact scan on every shard
            _same(idx.knn_batch(q[:1], k), tuple(a[:1] for a in want[k]))    # fewer queries than shards
        idx.vec_set.set_flat_path("scan")
        _same(idx.knn_batch(q[:40], 10), tuple(a[:40] for a in want[10]))
        idx.vec_set.set_flat_path(None)
        # the completeness check's fallback: every 3rd query is forced through the exact re-scan of the flagged queries
        before = V.lib().vdb_flat_gemm_fallbacks()
        V._lib.check(V.lib().vdb_debug_force_redo(3))
        _same(idx.knn_batch(q, 10), want[10])
        V._lib.check(V.lib().vdb_debug_force_redo(0))
        assert V.lib().vdb_flat_gemm_fallbacks() - before >= (nq + 2) // 3
        del idx
    # oracle parity of the sharded result (a query sample keeps the CPU time in seconds)
    ns = 24
    ores = oracle.flat_knn(base, q[:ns], 10, metric, nthreads=8)
    assert_knn_parity(base, q[:ns], metric, tuple(a[:ns] for a in want[10]), ores, oracle)


def test_uneven_device_resident_shards(vdb):
    """vdb_dataset_create_sharded_dev + vdb_flat_knn_sharded_dev: three uneven row blocks adopted in place, the batch
    resident on every shard's device, each shard receives the slice of the results it owns."""
    import torch
    V = vdb
    n, nq, dim, k = 300_000, 1000, 96, 100
    base, q = _data(n, nq, dim, 2)
    V.init_devices([])
    full = V.FlatIndex.from_vec_set(base, "l2sqr")
    full.vec_set.set_flat_path("scan")
    want = full.knn_batch(q, k)
    for devs in _device_sets():
        g = len(devs)
        counts = [70_000, 131_072] + [0] * (g - 2)
        rest = n - sum(counts)
        for s in range(2, g):
            counts[s] = rest // (g - 2) + (rest % (g - 2) if s == g - 1 else 0)
        if g == 2:
            counts = [70_000, n - 70_000]
        blocks, at = [], 0
        for s, d in enumerate(devs):
            blocks.append(torch.from_numpy(base[at:at + counts[s]]).to(f"cuda:{d}"))
            at += counts[s]
        vs = V.DeviceVecSet.from_device_shards([b.data_ptr() for b in blocks], counts, devs, dim, dim, np.float32,
                                               "l2sqr", keepalive=blocks)
        idx = V.FlatIndex(vs)
        per = -(-nq // g)
        qd = [torch.from_numpy(q).to(f"cuda:{d}") for d in devs]
        ids = [torch.full((per, k), -1, dtype=torch.int64, device=f"cuda:{d}") for d in devs]
        dd = [torch.zeros((per, k), dtype=torch.float32, device=f"cuda:{d}") for d in devs]
        cnt = [torch.zeros((per,), dtype=torch.int32, device=f"cuda:{d}") for d in devs]
        for forced in (0, 7):
            V._lib.check(V.lib().vdb_debug_force_redo(forced))
            idx.knn_batch_sharded_dev([t.data_ptr() for t in qd], nq, k, [t.data_ptr() for t in ids],
                                      [t.data_ptr() for t in dd], [t.data_ptr() for t in cnt])
            got = (torch.cat([t.cpu() for t in ids])[:nq].numpy(), torch.cat([t.cpu() for t in dd])[:nq].numpy(),
                   torch.cat([t.cpu() for t in cnt])[:nq].numpy())
            _same(got, want)
        V._lib.check(V.lib().vdb_debug_force_redo(0))
        _same(idx.knn_batch(q, k), want)     # the host-pointer call on the same adopted shards
        del idx, vs


def test_tiny_sets_and_edge_cases(vdb):
    V = vdb
    rng = np.random.default_rng(3)
    base = rng.random((10, 12), dtype=np.float32)
    q = rng.random((4, 12), dtype=np.float32)
    V.init_devices([])
    want = V.FlatIndex.from_vec_set(base, "l2sqr").knn_batch(q, 7)
    want_all = V.FlatIndex.from_vec_set(base, "l2sqr").knn_batch(q, 16)      # k > N: N results
    V.init_devices([0, 0, 0, 0])
    idx = V.FlatIndex.from_vec_set(base, "l2sqr")                            # shards of 3, 3, 2, 2 rows
    _same(idx.knn_batch(q, 7), want)
    got = idx.knn_batch(q, 16)
    _same(got, want_all)
    assert (got[2] == 10).all() and (got[0][:, 10:] == np.iinfo(np.uint64).max).all()
    ids, dist, counts = idx.knn_batch(q, 0)
    assert ids.shape == (4, 0) and (counts == 0).all()
    pairs = idx.knn(q[0], 3)                                                 # the trait's single-query call
    assert [p.index for p in pairs] == want[0][0, :3].tolist()
    # fewer rows than shards: some shards are empty
    idx2 = V.FlatIndex.from_vec_set(base[:2], "l2sqr")
    got = idx2.knn_batch(q, 5)
    assert (got[2] == 2).all()
    # an entry point without a row-sharded implementation refuses the parent handle instead of touching NULL rows
    out = np.zeros(10, np.float32)
    rc = V.lib().vdb_row_cache(idx.vec_set._h, V._lib.ptr(out))
    assert rc == V._lib.EUNSUPPORTED
    with pytest.raises(V.VdbError):
        idx.vec_set.push(base[:1])


def test_concurrent_calls_on_a_sharded_handle(vdb):
    """The reference searches from many threads under a read lock (src/database/mod.rs:248-256)."""
    import threading
    V = vdb
    base, q = _data(140_000, 64, 64, 4)
    V.init_devices([])
    full = V.FlatIndex.from_vec_set(base, "l2sqr")
    want = full.knn_batch(q, 10)
    V.init_devices([0, 0])
    idx = V.FlatIndex.from_vec_set(base, "l2sqr")
    errs, outs = [], [None] * 6

    def work(i):
        try:
            for _ in range(3):
                outs[i] = (idx.knn_batch(q, 10), idx.knn_batch(q[i:i + 1], 10))
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(6)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for i, (a, b) in enumerate(outs):
        _same(a, want)
        _same(b, tuple(x[i:i + 1] for x in want))


def test_sharded_ivf_and_pq_equal_the_unsharded_calls(vdb, oracle):
    """vdb_ivf_create / vdb_ivf_knn and vdb_pq_create / vdb_pq_knn on a row-sharded set: assignment, lists,
    codes and search results are bit-identical to the unsharded calls (and the oracle's on a query sample). Runs on
    one GPU with three shards on device 0, and with one shard per GPU when more are visible."""
    V = vdb
    n, nq, dim = 200_000, 96, 96
    base, q = _data(n, nq, dim, 7)
    base[150_001] = base[5]
    rng = np.random.default_rng(1)
    cent = np.ascontiguousarray(base[rng.permutation(n)[:32]])
    m = 24
    books = np.concatenate([np.ascontiguousarray(base[100:116, lo:hi]).reshape(-1) for lo, hi in V.pq_groups(dim, m)])
    cfg = V.PQConfig(4, m, "l2sqr")
    V.init_devices([])
    full = V.FlatIndex.from_vec_set(base, "l2sqr")
    ivf_full = V.IVFIndex(full.vec_set, cent)
    pq_full = V.PQTable(full.vec_set, cfg, books)
    want_ivf = {p: ivf_full.knn_with_ef_batch(q, 10, p) for p in (1, 4, 32)}
    want_pq = {(k, ef): full.knn_pq_batch(q, k, ef, pq_full) for k, ef in ((10, 240), (10, 5), (3, 600))}
    lists_full = ivf_full.clusters
    for devs in _device_sets():
        V.init_devices(devs)
        vs = V.DeviceVecSet(base, "l2sqr")
        ivf = V.IVFIndex(vs, cent)
        assert (ivf.assignment == ivf_full.assignment).all()
        for a, b in zip(ivf.clusters, lists_full):
            assert (a == b).all()
        for p, w in want_ivf.items():
            _same(ivf.knn_with_ef_batch(q, 10, p), w)
        _same(ivf.knn_with_ef_batch(q[:1], 10, 4), tuple(a[:1] for a in want_ivf[4]))
        pq = V.PQTable(vs, cfg, books)
        assert (pq.encoded_vec_set == pq_full.encoded_vec_set).all()
        flat = V.FlatIndex(vs)
        for (k, ef), w in want_pq.items():
            _same(flat.knn_pq_batch(q, k, ef, pq), w)
        lut_a, lut_b = pq.create_lookup(q[0]), pq_full.create_lookup(q[0])
        assert all((np.asarray(x) == np.asarray(y)).all() for x, y in zip(lut_a, lut_b))
        del ivf, pq, flat, vs
    codes = pq_full.encoded_vec_set
    off, mem = oracle.ivf_lists(ivf_full.assignment, 32)
    oi = oracle.ivf_knn(base, cent, off, mem, q[:16], 10, 4, "l2sqr", nthreads=8)
    assert_knn_parity(base, q[:16], "l2sqr", tuple(a[:16] for a in want_ivf[4]), oi, oracle)
    op = oracle.flat_knn_pq(base, codes, books, m, 4, q[:16], 10, 240, "l2sqr", nthreads=8)
    assert_knn_parity(base, q[:16], "l2sqr", tuple(a[:16] for a in want_pq[(10, 240)]), op, oracle)


@pytest.mark.parametrize("metric", ["l2sqr", "cosine"])
def test_global_pruning_bound_changes_no_bit_and_gathers_fewer_rows(vdb, metric, monkeypatch):
    """Between the contraction and the rerank the shards exchange order statistics of their candidates' upper bounds and cut
    their lists against the bound these certify TOGETHER (csrc/flat_gemm.cu: tensor_filter_begin / _finish). A shard alone
    can only prune against its own k-th upper bound and holds fewer than k candidates once the set is split: with the
    exchange the same bits come back and fewer rows reach the exact rerank."""
    V = vdb
    n, nq, dim, k = 400_000, 500, 128, 100
    base, q = _data(n, nq, dim, 5)
    V.init_devices([])
    full = V.FlatIndex.from_vec_set(base, metric)
    full.vec_set.set_flat_path("scan")
    want = full.knn_batch(q, k)
    V.init_devices([0, 0, 0, 0])
    idx = V.FlatIndex.from_vec_set(base, metric)
    lib = V.lib()

    def cands():
        out = [C.c_uint64(0) for _ in range(3)]
        lib.vdb_flat_gemm_stats(*[C.byref(x) for x in out])
        return out[1].value, out[2].value

    gathered = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("VDB_MG_GLOBAL_PRUNE", flag)
        c0, f0 = cands()
        _same(idx.knn_batch(q, k), want)
        c1, f1 = cands()
        gathered[flag] = c1 - c0
        assert f1 - f0 == 0, "no query may fall back to the exact scan on this set"
    monkeypatch.delenv("VDB_MG_GLOBAL_PRUNE")
    assert gathered["1"] < gathered["0"], gathered
    assert gathered["1"] >= nq * k
    # forced re-scans go through the same phases
    V._lib.check(lib.vdb_debug_force_redo(5))
    _same(idx.knn_batch(q, k), want)
    V._lib.check(lib.vdb_debug_force_redo(0))
