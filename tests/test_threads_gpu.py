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
