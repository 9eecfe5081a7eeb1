"""Tensor-core Flat path (K2): contraction scores vs numpy, then end-to-end identity with the exact scan."""
import ctypes as C

import numpy as np
import pytest

from conftest import assert_knn_parity

pytestmark = pytest.mark.gpu


def _scores(V, base, q, stride, c):
    import torch
    from lab_1806_vec_db_b200 import _lib as L
    from lab_1806_vec_db_b200.sharded import unpack_keys
    vs = V.DeviceVecSet(base, "l2sqr")
    dq = torch.from_numpy(q).cuda()
    ns = base.shape[0] // stride
    out = torch.empty((q.shape[0], ns), dtype=torch.int64, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.check(L.lib().vdb_debug_gemm_scores_dev(vs._h, C.c_void_p(dq.data_ptr()), q.shape[0], stride, c,
                                              C.c_void_p(out.data_ptr()), st))
    torch.cuda.synchronize()
    s, idx = unpack_keys(out.cpu().numpy().astype(np.uint64))
    return s, idx


@pytest.mark.parametrize("n,dim,nq,stride", [(1024, 960, 128, 1), (5000, 960, 300, 1), (5000, 960, 130, 3),
                                             (777, 100, 5, 1), (4096, 32, 256, 2), (3000, 2052, 64, 1)])
def test_contraction_scores_match_numpy(n, dim, nq, stride):
    import lab_1806_vec_db_b200 as V
    rng = np.random.default_rng(n + dim)
    base = rng.random((n, dim), dtype=np.float32)
    q = rng.random((nq, dim), dtype=np.float32)
    for c in (0.0, 4.1e-3):
        s, idx = _scores(V, base, q, stride, c)
        rows = base[::stride][: n // stride].astype(np.float64)
        qq = q.astype(np.float64)
        xn2 = (rows * rows).sum(1)
        want = xn2[None, :] - 2.0 * qq @ rows.T - c * np.sqrt((qq * qq).sum(1))[:, None] * np.sqrt(xn2)[None, :]
        assert (idx == np.arange(n // stride)[None, :]).all()
        # TF32 operand rounding: |error| <= 2 * 2^-9 * ||q|| ||x|| (+ fp32 accumulation)
        bound = 4.2e-3 * np.sqrt((qq * qq).sum(1))[:, None] * np.sqrt(xn2)[None, :] + 1e-4
        err = np.abs(s.astype(np.float64) - want)
        assert (err <= bound).all(), float((err / bound).max())
        # and the scores are genuinely tensor-core results, not the exact fp32 values
        assert err.max() > 0


def _synthetic(n, nq, seed=0):
    rng = np.random.default_rng(seed)
    proto = rng.random((256, 960), dtype=np.float32) * 0.15
    base = (proto[rng.integers(0, 256, n)] + 0.02 * rng.standard_normal((n, 960), dtype=np.float32)).clip(0, 1)
    q = (proto[rng.integers(0, 256, nq)] + 0.02 * rng.standard_normal((nq, 960), dtype=np.float32)).clip(0, 1)
    return np.ascontiguousarray(base, np.float32), np.ascontiguousarray(q, np.float32)


@pytest.mark.parametrize("k", [10, 100])
@pytest.mark.parametrize("metric", ["l2sqr", "cosine"])
def test_tensor_path_identical_to_exact_scan(oracle, k, metric):
    """The tensor cores only prune: ids and distances must equal the exact scan's (and the oracle's)."""
    import lab_1806_vec_db_b200 as V
    from lab_1806_vec_db_b200 import _lib as L
    n, nq = 140_000, 300
    base, q = _synthetic(n, nq)
    idx = V.FlatIndex.from_vec_set(base, metric)
    lib = L.lib()
    try:
        L.check(lib.vdb_flat_set_path(1))
        scan = idx.knn_batch(q, k)
        L.check(lib.vdb_flat_set_path(2))
        fb0 = lib.vdb_flat_gemm_fallbacks()
        tens = idx.knn_batch(q, k)
        fallbacks = lib.vdb_flat_gemm_fallbacks() - fb0
    finally:
        L.check(lib.vdb_flat_set_path(0))
    assert (tens[2] == scan[2]).all()
    assert (tens[0] == scan[0]).all(), float((tens[0] == scan[0]).mean())
    assert (tens[1].view(np.uint32) == scan[1].view(np.uint32)).all()  # rerank uses the scan's summation order
    want = oracle.flat_knn(base, q[:16], k, metric, 8)
    assert_knn_parity(base, q[:16], metric, tuple(a[:16] for a in tens), want, oracle)
    assert fallbacks < nq // 2, f"{fallbacks} of {nq} queries fell back to the exact scan"


def test_tensor_path_adversarial_ties_and_duplicates(oracle):
    """Exact duplicates and a uniform cloud (no neighbourhood structure): still identical to the scan."""
    import lab_1806_vec_db_b200 as V
    from lab_1806_vec_db_b200 import _lib as L
    rng = np.random.default_rng(7)
    n, nq, k = 131_072 + 77, 140, 20
    base = rng.random((n, 960), dtype=np.float32)
    base[1000:1040] = base[5]          # 40 exact copies
    q = rng.random((nq, 960), dtype=np.float32)
    q[0] = base[5]
    idx = V.FlatIndex.from_vec_set(base, "l2sqr")
    lib = L.lib()
    try:
        L.check(lib.vdb_flat_set_path(1))
        scan = idx.knn_batch(q, k)
        L.check(lib.vdb_flat_set_path(2))
        tens = idx.knn_batch(q, k)
    finally:
        L.check(lib.vdb_flat_set_path(0))
    assert (tens[0] == scan[0]).all() and (tens[1].view(np.uint32) == scan[1].view(np.uint32)).all()
    assert tens[0][0, :3].tolist() == [5, 1000, 1001]


@pytest.mark.parametrize("n,dim,nq,k", [(70_000, 12, 40, 6), (66_000, 13, 13, 1), (80_000, 100, 300, 1024),
                                        (65_536, 960, 12, 10), (100_000, 36, 1000, 50),
                                        (70_000, 960, 128, 10), (70_000, 960, 129, 100)])
def test_auto_path_edge_shapes_match_scan(n, dim, nq, k):
    """Auto path (tensor pruning from 12 queries up) on odd shapes: tiny / ragged dims (zero-filled TMA boxes),
    k = 1 and k = 1024, the smallest supported shard, and both sides of the 128-query boundary between single CTAs
    (M = 128) and CTA pairs (M = 256). Must be bit-identical to the forced exact scan."""
    import lab_1806_vec_db_b200 as V
    from lab_1806_vec_db_b200 import _lib as L
    rng = np.random.default_rng(n + dim)
    base = rng.random((n, dim), dtype=np.float32)
    q = rng.random((nq, dim), dtype=np.float32)
    q[0] = base[17]
    idx = V.FlatIndex.from_vec_set(base, "l2sqr")
    lib = L.lib()
    try:
        L.check(lib.vdb_flat_set_path(1))
        scan = idx.knn_batch(q, k)
        L.check(lib.vdb_flat_set_path(0))
        q0 = C.c_uint64(0)
        lib.vdb_flat_gemm_stats(C.byref(q0), None, None)
        auto = idx.knn_batch(q, k)
        q1 = C.c_uint64(0)
        lib.vdb_flat_gemm_stats(C.byref(q1), None, None)
    finally:
        L.check(lib.vdb_flat_set_path(0))
    assert q1.value - q0.value == nq, "the auto path did not take the tensor route"
    assert (auto[2] == scan[2]).all()
    assert (auto[0] == scan[0]).all(), float((auto[0] == scan[0]).mean())
    assert (auto[1].view(np.uint32) == scan[1].view(np.uint32)).all()
    assert auto[0][0, 0] == 17 and auto[1][0, 0] == 0.0


def test_cosine_tensor_path_zero_vectors_and_signs(oracle):
    """Cosine through the tensor route with signed data, exact zero rows and a zero query (the reference clamps the
    norm product at 1e-10, distance/mod.rs:67-69): identical to the forced scan and to the oracle."""
    import lab_1806_vec_db_b200 as V
    from lab_1806_vec_db_b200 import _lib as L
    rng = np.random.default_rng(3)
    base = rng.standard_normal((70_000, 64)).astype(np.float32)
    base[100:110] = 0.0
    base[500] *= 1e-7
    q = rng.standard_normal((20, 64)).astype(np.float32)
    q[3] = 0.0
    q[4] = base[1234] * 3.0
    idx = V.FlatIndex.from_vec_set(base, "cosine")
    lib = L.lib()
    try:
        L.check(lib.vdb_flat_set_path(1))
        scan = idx.knn_batch(q, 12)
        L.check(lib.vdb_flat_set_path(2))
        tens = idx.knn_batch(q, 12)
    finally:
        L.check(lib.vdb_flat_set_path(0))
    assert (tens[0] == scan[0]).all() and (tens[1].view(np.uint32) == scan[1].view(np.uint32)).all()
    want = oracle.flat_knn(base, q, 12, "cosine", 8)
    assert_knn_parity(base, q, "cosine", tens, want, oracle)
    assert tens[0][4, 0] == 1234


@pytest.mark.parametrize("metric", ["l2sqr", "cosine"])
def test_u8_rows_through_the_tensor_path(oracle, metric):
    """u8 rows and queries are exact in TF32, so the pruning bound only carries the accumulation term; results must
    be bit-identical to the forced scan (16-element steps of the u8 scan kernel) and match the oracle."""
    import lab_1806_vec_db_b200 as V
    from lab_1806_vec_db_b200 import _lib as L
    rng = np.random.default_rng(8)
    proto = rng.integers(0, 200, (64, 100), dtype=np.uint8)
    base = np.clip(proto[rng.integers(0, 64, 70_000)].astype(np.int16) + rng.integers(-20, 21, (70_000, 100)), 0, 255).astype(np.uint8)
    q = np.clip(proto[rng.integers(0, 64, 150)].astype(np.int16) + rng.integers(-20, 21, (150, 100)), 0, 255).astype(np.uint8)
    base[900:910] = base[3]
    idx = V.FlatIndex.from_vec_set(base, metric)
    lib = L.lib()
    try:
        L.check(lib.vdb_flat_set_path(1))
        scan = idx.knn_batch(q, 20)
        L.check(lib.vdb_flat_set_path(2))
        tens = idx.knn_batch(q, 20)
    finally:
        L.check(lib.vdb_flat_set_path(0))
    assert (tens[0] == scan[0]).all() and (tens[1].view(np.uint32) == scan[1].view(np.uint32)).all()
    want = oracle.flat_knn(base, q[:24], 20, metric, 8)
    assert_knn_parity(base, q[:24], metric, tuple(a[:24] for a in tens), want, oracle)


def test_auto_path_keeps_scan_for_small_shards(oracle):
    """Shards under 65536 rows have no tensor route: the auto path must silently use the exact scan."""
    import lab_1806_vec_db_b200 as V
    rng = np.random.default_rng(3)
    b8 = rng.integers(0, 256, (7_000, 32), dtype=np.uint8)
    q8 = rng.integers(0, 256, (20, 32), dtype=np.uint8)
    got = V.FlatIndex.from_vec_set(b8, "l2sqr").knn_batch(q8, 5)
    want = oracle.flat_knn(b8, q8, 5, "l2sqr", 8)
    assert_knn_parity(b8, q8, "l2sqr", got, want, oracle)


def test_tensor_path_large_query_batch_is_chunked():
    """40 000 queries (three chunks of the tensor path) against a small shard: identical to the scan on a subsample."""
    import lab_1806_vec_db_b200 as V
    from lab_1806_vec_db_b200 import _lib as L
    rng = np.random.default_rng(11)
    base = rng.random((66_000, 32), dtype=np.float32)
    q = rng.random((40_000, 32), dtype=np.float32)
    idx = V.FlatIndex.from_vec_set(base, "l2sqr")
    auto = idx.knn_batch(q, 5)
    sel = np.r_[0:50, 16380:16390, 32760:32780, 39990:40000]
    lib = L.lib()
    try:
        L.check(lib.vdb_flat_set_path(1))
        scan = idx.knn_batch(q[sel], 5)
    finally:
        L.check(lib.vdb_flat_set_path(0))
    assert (auto[0][sel] == scan[0]).all() and (auto[1][sel].view(np.uint32) == scan[1].view(np.uint32)).all()
