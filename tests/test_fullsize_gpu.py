"""BASELINE.json configs[1] at its FULL size (1M x 960 f32 rows, 10 000-query batch, k = 100): the oracle cannot cover
10^10 pairs, so the batch is checked through size-independent properties — planted exact copies, sortedness by
(distance, id), idempotence, invariance under the code path (tensor pruning with / without row parts, exact streaming
scan) and under row sharding (3 uneven shards merged by key), all BIT-exact — plus an oracle parity check on a query
sample (the rows are bench.py's synthetic GIST-shaped set)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import assert_knn_parity

pytestmark = pytest.mark.gpu

N, NQ, K = 1_000_000, 10_000, 100


def _dev_out(torch, nq, k, dev):
    return (torch.empty((nq, k), dtype=torch.int64, device=dev), torch.empty((nq, k), dtype=torch.float32, device=dev),
            torch.empty((nq,), dtype=torch.int32, device=dev))


def test_full_size_flat_batch_properties(oracle):
    import torch
    import lab_1806_vec_db_b200 as V
    from lab_1806_vec_db_b200 import _lib as L
    from bench import DIM, load_fixtures, synth

    dev = torch.device("cuda:0")
    lib = L.lib()
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    b1000, t1000 = load_fixtures()
    base = synth(b1000, 0, N, 42, dev)
    q = synth(t1000, 0, NQ, 43, dev)
    rng = np.random.default_rng(11)
    planted = torch.as_tensor(rng.choice(N, 64, replace=False), device=dev)
    q[:64] = base[planted]                                   # exact copies: rank 0, distance 0
    vs = V.DeviceVecSet.from_device(base.data_ptr(), N, DIM, DIM, np.float32, "l2sqr", keepalive=base)

    def flat(queries, k=K):
        ids, dd, cnt = _dev_out(torch, queries.shape[0], k, dev)
        L.check(lib.vdb_flat_knn_dev(vs._h, C.c_void_p(queries.data_ptr()), queries.shape[0], k, C.c_void_p(ids.data_ptr()),
                                     C.c_void_p(dd.data_ptr()), C.c_void_p(cnt.data_ptr()), st))
        torch.cuda.synchronize()
        return ids, dd, cnt

    def same(a, b):
        return bool((a[0] == b[0]).all() and (a[1].view(torch.int32) == b[1].view(torch.int32)).all() and (a[2] == b[2]).all())

    keep = os.environ.get("VDB_GEMM_PARTS")
    try:
        q0 = C.c_uint64(0)
        lib.vdb_flat_gemm_stats(C.byref(q0), None, None)
        res = flat(q)                                        # auto path: tensor pruning, filter pass in row parts
        q1 = C.c_uint64(0)
        lib.vdb_flat_gemm_stats(C.byref(q1), None, None)
        assert q1.value - q0.value == NQ, "the 10 000-query batch did not take the tensor route"
        ids, dd, cnt = res
        # ---- properties ----
        assert bool((cnt == K).all())
        assert bool((ids[:64, 0] == planted).all()) and bool((dd[:64, 0] == 0).all())
        dif = dd[:, 1:] - dd[:, :-1]
        assert bool((dif >= 0).all()), "distances not ascending"
        tie = dif == 0
        assert bool((ids[:, 1:][tie] > ids[:, :-1][tie]).all()), "ties not ordered by id"
        srt = torch.sort(ids, dim=1).values
        assert bool((srt[:, 1:] != srt[:, :-1]).all()), "duplicate ids in a result"
        assert bool((ids >= 0).all() and (ids < N).all())
        # ---- idempotence ----
        assert same(res, flat(q))
        # ---- code-path invariance (bit-exact) ----
        os.environ["VDB_GEMM_PARTS"] = "1"
        assert same(res, flat(q)), "row parts changed the result"
        os.environ["VDB_GEMM_PARTS"] = "5"
        assert same(res, flat(q)), "5 row parts changed the result"
        os.environ.pop("VDB_GEMM_PARTS")
        sub = torch.cat([torch.arange(0, 24, device=dev), torch.as_tensor(rng.choice(NQ, 40, replace=False), device=dev)])
        qs = q.index_select(0, sub).contiguous()
        L.check(lib.vdb_flat_set_path(1))                    # exact streaming scan
        scan = flat(qs)
        L.check(lib.vdb_flat_set_path(0))
        assert same(tuple(t.index_select(0, sub) for t in res), scan), "tensor route and streaming scan differ"
        small = flat(qs[:7])                                 # 7 queries: single-CTA tensor pass
        assert same(tuple(t[:7] for t in scan), small)
        # ---- row-sharding invariance: 3 uneven shards, per-shard keys merged by (distance, global id) ----
        bounds = [0, 300_017, 650_001, N]
        keys = torch.empty((3, NQ, K), dtype=torch.int64, device=dev)
        shards = []
        for s in range(3):
            lo, hi = bounds[s], bounds[s + 1]
            part = base[lo:hi]
            sh = V.DeviceVecSet.from_device(part.data_ptr(), hi - lo, DIM, DIM, np.float32, "l2sqr", id_base=lo, keepalive=base)
            shards.append(sh)
            L.check(lib.vdb_flat_knn_keys_dev(sh._h, C.c_void_p(q.data_ptr()), NQ, K, C.c_void_p(keys[s].data_ptr()), st))
        merged = _dev_out(torch, NQ, K, dev)
        L.check(lib.vdb_merge_keys_dev(C.c_void_p(keys.data_ptr()), 3, NQ, K, C.c_void_p(merged[0].data_ptr()),
                                       C.c_void_p(merged[1].data_ptr()), C.c_void_p(merged[2].data_ptr()), st))
        torch.cuda.synchronize()
        assert same(res, merged), "sharded + merged result differs from the unsharded one"
        for sh in shards:
            sh.close()
        # ---- oracle parity on a query sample (8 x 10^6 pairs on the CPU) ----
        base_h = base.cpu().numpy()
        pick = [0, 1, 63, 64, 65, 4999, 9998, 9999]
        qh = q.cpu().numpy()[pick]
        want = oracle.flat_knn(base_h, qh, K, "l2sqr", os.cpu_count() or 1)
        got = tuple(t.cpu().numpy()[pick] for t in res)
        rate = assert_knn_parity(base_h, qh, "l2sqr", (got[0].astype(np.uint64), got[1], got[2].astype(np.uint32)), want, oracle)
        assert rate > 0.97
    finally:
        L.check(lib.vdb_flat_set_path(0))
        if keep is None:
            os.environ.pop("VDB_GEMM_PARTS", None)
        else:
            os.environ["VDB_GEMM_PARTS"] = keep
        vs.close()


def test_full_size_ivf_and_pq_properties(oracle):
    """configs[2] / configs[3] at 1M x 960: IVF assignment bit-exact on a row sample, probing EVERY list returns the exact
    Flat result bit for bit (the reference's cross-index test, ivf_index.rs:222-232, at full size), nprobe = 8 equals the
    oracle's probe scan; PQ codes bit-exact on a row sample and knn_pq ids equal the oracle's on the same codebooks."""
    import torch
    import lab_1806_vec_db_b200 as V
    from lab_1806_vec_db_b200 import _lib as L
    from lab_1806_vec_db_b200.index import train_codebooks
    from bench import DIM, load_fixtures, synth

    dev = torch.device("cuda:0")
    lib = L.lib()
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    b1000, t1000 = load_fixtures()
    base = synth(b1000, 0, N, 42, dev)
    nq, k = 256, 10
    q = synth(t1000, 0, nq, 43, dev)
    vs = V.DeviceVecSet.from_device(base.data_ptr(), N, DIM, DIM, np.float32, "l2sqr", keepalive=base)
    base_h = base.cpu().numpy()
    q_h = q.cpu().numpy()
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(42)
    try:
        ids, dd, cnt = _dev_out(torch, nq, k, dev)
        L.check(lib.vdb_flat_knn_dev(vs._h, C.c_void_p(q.data_ptr()), nq, k, C.c_void_p(ids.data_ptr()),
                                     C.c_void_p(dd.data_ptr()), C.c_void_p(cnt.data_ptr()), st))
        torch.cuda.synchronize()
        flat = (ids.clone(), dd.clone(), cnt.clone())
        # ---- IVF ----
        train = np.ascontiguousarray(base_h[rng.permutation(N)[:100_000]])
        km = V.KMeans.from_vec_set(train, V.KMeansConfig(128, 20, 1e-6, "l2sqr"), rng)
        ivf = V.IVFIndex(vs, km.centroids)
        ns = 20_000
        assert (oracle.kmeans_assign(base_h[:ns], km.centroids, "l2sqr", nthreads=cores) == ivf.assignment[:ns]).all()
        assert sum(len(c) for c in ivf.clusters) == N

        def ivf_search(nprobe):
            L.check(lib.vdb_ivf_knn_dev(vs._h, ivf._h, C.c_void_p(q.data_ptr()), nq, k, nprobe, C.c_void_p(ids.data_ptr()),
                                        C.c_void_p(dd.data_ptr()), C.c_void_p(cnt.data_ptr()), st))
            torch.cuda.synchronize()
            return ids.clone(), dd.clone(), cnt.clone()
        full = ivf_search(128)
        assert bool((full[0] == flat[0]).all()), float((full[0] == flat[0]).float().mean())
        assert bool((full[1].view(torch.int32) == flat[1].view(torch.int32)).all())
        part = ivf_search(8)
        off, mem = oracle.ivf_lists(ivf.assignment, 128)
        nc = 6
        oi, od, oc = oracle.ivf_knn(base_h, km.centroids, off, mem, q_h[:nc], k, 8, "l2sqr", nthreads=cores)
        got = (part[0][:nc].cpu().numpy().astype(np.uint64), part[1][:nc].cpu().numpy(), part[2][:nc].cpu().numpy().astype(np.uint32))
        assert_knn_parity(base_h, q_h[:nc], "l2sqr", got, (oi, od, oc), oracle)
        ivf.close()
        # ---- PQ ----
        cfg = V.PQConfig(4, 240, "l2sqr", 10_000, 20, 1e-6)
        tdev = V.DeviceVecSet(np.ascontiguousarray(base_h[rng.permutation(N)[:10_000]]), "l2sqr")
        books = train_codebooks(tdev, cfg, rng)
        tdev.close()
        pq = V.PQTable(vs, cfg, books)
        assert (oracle.pq_encode(base_h[:2000], books, 240, 4, "l2sqr", nthreads=1) == pq.encoded_vec_set[:2000]).all()
        ef = 240
        L.check(lib.vdb_pq_knn_dev(vs._h, pq._h, C.c_void_p(q.data_ptr()), nq, k, ef, C.c_void_p(ids.data_ptr()),
                                   C.c_void_p(dd.data_ptr()), C.c_void_p(cnt.data_ptr()), st))
        torch.cuda.synchronize()
        nc = 4
        oi, od, oc = oracle.flat_knn_pq(base_h, pq.encoded_vec_set, books, 240, 4, q_h[:nc], k, ef, "l2sqr", nthreads=cores)
        got = (ids[:nc].cpu().numpy().astype(np.uint64), dd[:nc].cpu().numpy(), cnt[:nc].cpu().numpy().astype(np.uint32))
        assert_knn_parity(base_h, q_h[:nc], "l2sqr", got, (oi, od, oc), oracle)
        pq.close()
    finally:
        vs.close()
