"""Tensor-core Flat path (K2): contraction scores vs numpy, then end-to-end identity with the exact scan."""
import ctypes as C

import numpy as np
import pytest

from conftest import assert_knn_parity

pytestmark = pytest.mark.gpu


def _scores(V, base, q, stride, kind, metric="l2sqr"):
    import torch
    from lab_1806_vec_db_b200 import _lib as L
    from lab_1806_vec_db_b200.sharded import unpack_keys
    vs = V.DeviceVecSet(base, metric)
    dq = torch.from_numpy(q).cuda()
    ns = base.shape[0] // stride
    out = torch.empty((q.shape[0], ns), dtype=torch.int64, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.check(L.lib().vdb_debug_gemm_scores_dev(vs._h, C.c_void_p(dq.data_ptr()), q.shape[0], stride, kind,
                                              C.c_void_p(out.data_ptr()), st))
    torch.cuda.synchronize()
    s, idx = unpack_keys(out.cpu().numpy().astype(np.uint64))
    return s, idx


def _operand_errors(x, kind):
    """||x - x~|| per row for the operand the tensor core consumes (numpy model of csrc/flat_gemm.cu::row_side_kernel):
    kind 1: x~ = fp16(x * s) / s with one power-of-two scale for the whole array; kind 0: fp32 bits truncated to TF32."""
    x = np.asarray(x, np.float32)
    if kind == 1:
        m = float(np.abs(x).max())
        s = np.float32(2.0 ** (14 - np.frexp(m)[1])) if m > 0 else np.float32(1)
        back = (x * s).astype(np.float16).astype(np.float32) / s
    else:
        back = (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    return np.sqrt(((x.astype(np.float64) - back.astype(np.float64)) ** 2).sum(1))


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("n,dim,nq,stride", [(1024, 960, 128, 1), (5000, 960, 300, 1), (5000, 960, 130, 3),
                                             (777, 100, 5, 1), (4096, 32, 256, 2), (3000, 2052, 64, 1)])
def test_pruning_scores_are_tight_lower_bounds(n, dim, nq, stride, kind):
    """The invariant the exactness of the tensor path rests on: for EVERY (query, row) pair the pruning score S' is a
    lower bound of d(q, x) - ||q||^2, and it is no further below it than twice the rigorous bound
    b = 2 (||q|| ex + eq ||x|| + dim 2^-23 ||q|| ||x||) built from the measured operand errors ex, eq."""
    import lab_1806_vec_db_b200 as V
    rng = np.random.default_rng(n + dim)
    base = rng.random((n, dim), dtype=np.float32)
    base[::7] *= np.float32(1e-3)                      # rows of very different magnitude share one FP16 scale
    q = rng.random((nq, dim), dtype=np.float32) * np.float32(3.0)
    s, idx = _scores(V, base, q, stride, kind)
    rows = base[::stride][: n // stride]
    assert (idx == np.arange(n // stride)[None, :]).all()
    r64, q64 = rows.astype(np.float64), q.astype(np.float64)
    xn, qn = np.sqrt((r64 * r64).sum(1)), np.sqrt((q64 * q64).sum(1))
    exact = (xn * xn)[None, :] - 2.0 * q64 @ r64.T
    ex = _operand_errors(base, kind)[::stride][: n // stride]
    if kind == 1:   # queries carry their own power-of-two scale
        eq = np.array([_operand_errors(q[i:i + 1], 1)[0] for i in range(nq)])
    else:           # queries are rounded to nearest TF32: at most half the truncation error per element
        eq = _operand_errors(q, 0)
    b = 2.0 * (qn[:, None] * ex[None, :] + eq[:, None] * xn[None, :] + dim * 2.0 ** -23 * qn[:, None] * xn[None, :])
    gap = exact - s.astype(np.float64)
    fp32_noise = 4e-6 * ((xn * xn)[None, :] + qn[:, None] * xn[None, :])
    assert (gap >= -fp32_noise).all(), float((gap / fp32_noise).min())          # lower bound
    assert (gap <= 2.0 * 1.01 * b + fp32_noise).all(), float((gap / (2 * b + fp32_noise)).max())   # and a tight one
    # the scores are genuinely tensor-core results of rounded operands, not exact fp32 values minus the bound
    assert np.abs(gap - b).max() > 0


def test_cosine_pruning_scores_are_lower_bounds():
    import lab_1806_vec_db_b200 as V
    rng = np.random.default_rng(5)
    base = rng.random((3000, 480), dtype=np.float32)
    base[17] = 0.0                                       # a zero row sits under the reference's 1e-10 clamp: always kept
    q = rng.random((200, 480), dtype=np.float32)
    for kind in (0, 1):
        s, _ = _scores(V, base, q, 1, kind, "cosine")
        r64, q64 = base.astype(np.float64), q.astype(np.float64)
        xn, qn = np.sqrt((r64 * r64).sum(1)), np.sqrt((q64 * q64).sum(1))
        exact = 1.0 - (q64 @ r64.T) / np.maximum(qn[:, None] * xn[None, :], 1e-10)
        assert np.isneginf(s[:, 17]).all()
        keep = np.arange(3000) != 17
        gap = exact[:, keep] - s[:, keep].astype(np.float64)
        assert (gap >= -2e-6).all(), float(gap.min())
        assert (gap <= 4.5e-3).all(), float(gap.max())


def _synthetic(n, nq, seed=0):
    rng = np.random.default_rng(seed)
    proto = rng.random((256, 960), dtype=np.float32) * 0.15
    base = (proto[rng.integers(0, 256, n)] + 0.02 * rng.standard_normal((n, 960), dtype=np.float32)).clip(0, 1)
    q = (proto[rng.integers(0, 256, nq)] + 0.02 * rng.standard_normal((nq, 960), dtype=np.float32)).clip(0, 1)
    return np.ascontiguousarray(base, np.float32), np.ascontiguousarray(q, np.float32)


@pytest.mark.parametrize("k", [10, 100])
@pytest.mark.parametrize("metric", ["l2sqr", "cosine"])
@pytest.mark.parametrize("kind", ["f16", "tf32"])
def test_tensor_path_identical_to_exact_scan(oracle, k, metric, kind, monkeypatch):
    """The tensor cores only prune: ids and distances must equal the exact scan's (and the oracle's), whichever
    operand kind feeds the contraction (FP16 copy, or the fp32 rows truncated to TF32 in place)."""
    import lab_1806_vec_db_b200 as V
    from lab_1806_vec_db_b200 import _lib as L
    monkeypatch.setenv("VDB_GEMM_KIND", kind)
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


def test_nonfinite_rows_and_queries_take_the_exact_route(oracle):
    """Rows with inf / NaN components are always kept as candidates (the exact rerank gives them their inf / NaN
    distance, which sorts last), a NaN query is answered by the exact scan: identical to the forced scan."""
    import lab_1806_vec_db_b200 as V
    from lab_1806_vec_db_b200 import _lib as L
    base, q = _synthetic(70_000, 40, seed=11)
    base[5, 7] = np.inf
    base[6, 0] = np.nan
    q[3, 100] = np.nan
    idx = V.FlatIndex.from_vec_set(base, "l2sqr")
    lib = L.lib()
    try:
        L.check(lib.vdb_flat_set_path(1))
        scan = idx.knn_batch(q, 10)
        L.check(lib.vdb_flat_set_path(2))
        tens = idx.knn_batch(q, 10)
    finally:
        L.check(lib.vdb_flat_set_path(0))
    assert (tens[0] == scan[0]).all() and (tens[1].view(np.uint32) == scan[1].view(np.uint32)).all()
    assert np.isnan(tens[1][3]).all() and not np.isnan(np.delete(tens[1], 3, 0)).any()


def test_rows_of_wildly_different_magnitude_fall_back_to_tf32_operands():
    """Cosine over rows whose magnitudes span 12 decades: one FP16 scale for the set flushes the small rows, the
    measured operand-error norm shows it and the TF32 kind (fp32 rows in place) is selected; results still equal the scan."""
    import lab_1806_vec_db_b200 as V
    from lab_1806_vec_db_b200 import _lib as L
    rng = np.random.default_rng(12)
    base = rng.standard_normal((70_000, 64)).astype(np.float32) * (10.0 ** rng.uniform(-6, 6, (70_000, 1))).astype(np.float32)
    q = rng.standard_normal((64, 64)).astype(np.float32)
    q[:32] = base[rng.integers(0, 70_000, 32)] * np.float32(1.5) + np.float32(0.01) * q[:32]
    idx = V.FlatIndex.from_vec_set(base, "cosine")
    lib = L.lib()
    try:
        L.check(lib.vdb_flat_set_path(1))
        scan = idx.knn_batch(q, 10)
        L.check(lib.vdb_flat_set_path(2))
        fb0 = lib.vdb_flat_gemm_fallbacks()
        tens = idx.knn_batch(q, 10)
        fallbacks = lib.vdb_flat_gemm_fallbacks() - fb0
    finally:
        L.check(lib.vdb_flat_set_path(0))
    assert (tens[0] == scan[0]).all() and (tens[1].view(np.uint32) == scan[1].view(np.uint32)).all()
    assert fallbacks < 32


@pytest.mark.parametrize("metric", ["l2sqr", "cosine"])
def test_round2_stages_do_not_change_a_bit(oracle, metric, monkeypatch):
    """Upper-bound pruning of the candidate lists, the two-level sample selection, the filter's two rare paths and the row
    parts only change how many rows reach the exact rerank: every combination returns the bits of the exact scan, the
    pruned run gathers fewer rows, and nothing falls back to the scan on this set."""
    import lab_1806_vec_db_b200 as V
    from lab_1806_vec_db_b200 import _lib as L
    n, nq, k = 150_000, 400, 50
    base, q = _synthetic(n, nq)
    idx = V.FlatIndex.from_vec_set(base, metric)
    lib = L.lib()
    stats = lambda: [x.value for x in _stats(lib)]  # noqa: E731
    try:
        L.check(lib.vdb_flat_set_path(1))
        scan = idx.knn_batch(q, k)
        L.check(lib.vdb_flat_set_path(2))
        runs = {}
        for name, env in (("default", {}), ("no_prune", {"VDB_GEMM_PRUNE": "0"}), ("one_level", {"VDB_GEMM_SAMPLE_2L": "0"}), ("two_level", {"VDB_GEMM_SAMPLE_2L": "1"}),
                          ("rare_mask", {"VDB_GEMM_RARE_PER_SCORE": "0"}), ("rare_per_score", {"VDB_GEMM_RARE_PER_SCORE": "1"}),
                          ("rare_staged", {"VDB_GEMM_RARE_PER_SCORE": "2"}),
                          ("parts3", {"VDB_GEMM_PARTS": "3"})):
            for key, val in env.items():
                monkeypatch.setenv(key, val)
            s0 = stats()
            runs[name] = (idx.knn_batch(q, k), [b - a for a, b in zip(s0, stats())])
            for key in env:
                monkeypatch.delenv(key)
    finally:
        L.check(lib.vdb_flat_set_path(0))
    for name, (res, (queries, cands, fallbacks)) in runs.items():
        assert queries == nq, name
        assert (res[0] == scan[0]).all(), name
        assert (res[1].view(np.uint32) == scan[1].view(np.uint32)).all(), name
        assert (res[2] == scan[2]).all(), name
        assert fallbacks == 0, (name, fallbacks)
    assert runs["default"][1][1] < runs["no_prune"][1][1], "pruning did not reduce the rows gathered by the rerank"
    assert runs["default"][1][1] >= nq * k
    want = oracle.flat_knn(base, q[:16], k, metric, 8)
    assert_knn_parity(base, q[:16], metric, tuple(a[:16] for a in runs["default"][0]), want, oracle)


def _stats(lib):
    out = [C.c_uint64(0) for _ in range(3)]
    lib.vdb_flat_gemm_stats(*[C.byref(x) for x in out])
    return out


def test_pruning_keeps_rows_under_the_cosine_norm_clamp(oracle):
    """Rows whose norm product falls under the reference's 1e-10 clamp carry the score -inf ("always a candidate"); the
    upper-bound pruning must not count them as rows with a known small distance, nor drop them."""
    import lab_1806_vec_db_b200 as V
    from lab_1806_vec_db_b200 import _lib as L
    rng = np.random.default_rng(5)
    n, nq, k = 70_000, 200, 20
    base = rng.random((n, 64), dtype=np.float32)
    base[rng.choice(n, 300, replace=False)] = 0.0          # zero rows: cosine distance 1 - 0 / 1e-10 = 1
    base[rng.choice(n, 300, replace=False)] *= 1e-12        # tiny rows: under the clamp for every query
    q = rng.random((nq, 64), dtype=np.float32)
    idx = V.FlatIndex.from_vec_set(base, "cosine")
    lib = L.lib()
    try:
        L.check(lib.vdb_flat_set_path(1))
        scan = idx.knn_batch(q, k)
        L.check(lib.vdb_flat_set_path(2))
        tens = idx.knn_batch(q, k)
    finally:
        L.check(lib.vdb_flat_set_path(0))
    assert (tens[0] == scan[0]).all() and (tens[1].view(np.uint32) == scan[1].view(np.uint32)).all()
