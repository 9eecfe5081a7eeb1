"""Re-entrancy of the search entry points on one handle (the reference searches under a read lock from rayon workers
and Python threads: database/mod.rs:255, examples/test_multi_threads.py)."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_concurrent_searches_on_one_handle(fixtures):
    import lab_1806_vec_db_b200 as V
    rng = np.random.default_rng(0)
    base = rng.random((80_000, 64), dtype=np.float32)
    queries = rng.random((8, 40, 64), dtype=np.float32)
    flat = V.FlatIndex.from_vec_set(base, "l2sqr")
    cent = np.ascontiguousarray(base[:16])
    ivf = V.IVFIndex(flat.vec_set, cent)
    books = np.concatenate([np.ascontiguousarray(base[100:116, lo:hi]).reshape(-1) for lo, hi in V.pq_groups(64, 16)])
    pq = V.PQTable(flat.vec_set, V.PQConfig(4, 16, "l2sqr"), books)          # batches of 40: tensor-core ADC filter
    hnsw = V.HNSWIndex(flat.vec_set, V.HNSWConfig(0, 40, 8), rng)

    def all_searches(q):
        return (flat.knn_batch(q, 7), flat.knn_batch(q[:3], 7), ivf.knn_with_ef_batch(q, 7, 4),
                flat.knn_pq_batch(q, 7, 60, pq), hnsw.knn_with_ef_batch(q, 7, 30), hnsw.knn_pq_batch(q, 7, 30, pq))

    serial = [all_searches(q) for q in queries]
    out, errs = [None] * 8, []

    def work(i):
        try:
            for _ in range(3):
                out[i] = all_searches(queries[i])
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for got, want in zip(out, serial):
        for g, w in zip(got, want):
            assert (g[0] == w[0]).all() and (g[1].view(np.uint32) == w[1].view(np.uint32)).all() and (g[2] == w[2]).all()


def test_concurrent_single_query_calls_are_coalesced(fixtures):
    """The reference's call pattern: one query per knn call from many threads (examples/bench.rs:410-416,
    src/database/mod.rs:248-256). Concurrent calls share database passes; every caller still gets exactly the bits
    of its own individual call."""
    import ctypes as C
    import lab_1806_vec_db_b200 as V
    from lab_1806_vec_db_b200 import _lib as L
    rng = np.random.default_rng(1)
    base = rng.random((200_000, 96), dtype=np.float32)
    queries = rng.random((8, 25, 96), dtype=np.float32)
    flat = V.FlatIndex.from_vec_set(base, "l2sqr")
    lib = L.lib()
    L.check(lib.vdb_set_batching(0))
    try:
        serial = [[flat.knn_batch(q[i:i + 1], 10) for i in range(q.shape[0])] for q in queries]
    finally:
        L.check(lib.vdb_set_batching(1))
    out, errs = [[None] * 25 for _ in range(8)], []

    def work(t):
        try:
            for i in range(25):
                out[t][i] = flat.knn_batch(queries[t][i:i + 1], 10 if (t + i) % 5 else 7)   # mixed k in flight
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for t in range(8):
        for i in range(25):
            k = 10 if (t + i) % 5 else 7
            g, w = out[t][i], serial[t][i]
            assert (g[0] == w[0][:, :k]).all() and (g[1].view(np.uint32) == w[1][:, :k].view(np.uint32)).all()
            assert (g[2] == k).all()
    nb, nq = C.c_uint64(0), C.c_uint64(0)
    L.check(lib.vdb_batch_stats(flat.vec_set._h, C.byref(nb), C.byref(nq)))
    assert nq.value == 200 and nb.value < 200, (nb.value, nq.value)   # passes were shared
    # the trait's single-query call through the same path
    assert [p.index for p in flat.knn(queries[0][0], 5)] == serial[0][0][0][0, :5].tolist()
