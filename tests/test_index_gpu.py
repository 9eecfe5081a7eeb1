"""Parity of the CUDA k-means / PQ / IVF / distance paths (through the C ABI) against the CPU oracle.

Index-valued results (assignments, codes, lists, probe order) and the quantities they are derived from
(lookup tables, ADC sums) must be BIT-EXACT; reranked distances follow the 1e-5 rule."""
import numpy as np
import pytest

from conftest import ATOL, RTOL, assert_knn_parity, close

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def V():
    import lab_1806_vec_db_b200 as V
    return V


# ---- distance primitives ------------------------------------------------------------------------------
def test_calc_dist_known_answers(V):
    """distance/mod.rs:138-150 and the pyo3 calc_dist default (cosine)."""
    assert abs(V.calc_dist([1, 2, 3], [4, 5, 6], "l2sqr") - 27.0) < 1e-6
    assert abs(V.calc_dist([1, 2, 3], [2, 4, 6]) - 0.0) < 1e-6
    a8, b8 = np.array([[1, 2, 3]], np.uint8), np.array([[2, 4, 6]], np.uint8)
    assert abs(V.calc_dist_batch(a8, b8, "cosine")[0]) < 1e-6
    with pytest.raises(ValueError):
        V.calc_dist([1, 2], [1, 2, 3])
    with pytest.raises(ValueError):
        V.calc_dist([1, 2], [1, 2], "manhattan")


@pytest.mark.parametrize("metric", ["l2sqr", "cosine", "dot"])
def test_calc_dist_batch_vs_oracle(V, fixtures, oracle, metric):
    a, b = fixtures["base"][:300], fixtures["test"][:300]
    got = V.calc_dist_batch(a, b, metric)
    want = np.array([oracle.dot(x, y) if metric == "dot" else oracle.distance(x, y, metric) for x, y in zip(a, b)])
    assert close(got, want).all()


@pytest.mark.parametrize("metric", ["l2sqr", "cosine"])
def test_row_cache_and_gather_dist(V, fixtures, golden, oracle, metric):
    """K3/K10: dist_cache per row and cached-form candidate distances (hnsw_index.rs:351-358)."""
    base, test = fixtures["base"], fixtures["test"]
    vs = V.DeviceVecSet(base, metric)
    cache = V.dist_cache(vs)
    assert close(cache, golden[f"cached_{metric}_rowcache"]).all()
    cand = np.arange(0, 1000, 7)
    got = V.gather_dist(vs, test[3:5], [cand, cand[:5]])
    assert close(got[0], golden[f"cached_{metric}_dist"], RTOL, 2e-6).all()
    want1 = oracle.gather_dist(base, golden[f"cached_{metric}_rowcache"], test[4], oracle.dist_cache(test[4], metric),
                               cand[:5], metric)
    assert close(got[1], want1, RTOL, 2e-6).all()
    assert [len(g) for g in V.gather_dist(vs, test[:2], [[], [1]])] == [0, 1]
    with pytest.raises(V.VdbError):
        V.gather_dist(vs, test[:1], [[1000]])


# ---- k-means -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("metric", ["l2sqr", "cosine"])
def test_assign_bit_exact_golden(V, fixtures, golden, metric):
    base = fixtures["base"]
    km = V.KMeans(V.KMeansConfig(16, dist=metric), base[7:23])
    assert (km.find_nearest_batch(base) == golden[f"assign16_{metric}"]).all()


def test_assign_selected_range_golden(V, fixtures, golden):
    base = fixtures["base"]
    km = V.KMeans(V.KMeansConfig(16, dist="l2sqr", selected=(100, 113)), np.ascontiguousarray(base[7:23, 100:113]))
    assert (km.find_nearest_batch(base) == golden["assign16_sel_l2sqr"]).all()
    assert km.find_nearest(base[5]) == golden["assign16_sel_l2sqr"][5]


@pytest.mark.parametrize("k,d", [(1, 7), (3, 5), (16, 4), (33, 12), (128, 960), (200, 64), (5, 1500)])
@pytest.mark.parametrize("metric", ["l2sqr", "cosine"])
def test_assign_bit_exact_shapes(V, oracle, k, d, metric):
    rng = np.random.default_rng(k * 7 + d)
    n = 700 if d >= 960 else 3000
    rows = rng.random((n, d + 3), dtype=np.float32)
    cent = rng.random((k, d), dtype=np.float32)
    cent[k // 2] = cent[0]  # an exact duplicate centroid: ties must go to the lowest id
    km = V.KMeans(V.KMeansConfig(k, dist=metric, selected=(2, 2 + d)), cent)
    want = oracle.kmeans_assign(rows, cent, metric, sel=(2, 2 + d), nthreads=8)
    assert (km.find_nearest_batch(rows) == want).all()


def test_assign_u8_bit_exact(V, oracle):
    rng = np.random.default_rng(2)
    rows = rng.integers(0, 256, (2000, 48), dtype=np.uint8)
    cent = rng.integers(0, 256, (20, 48), dtype=np.uint8)
    for metric in ("l2sqr", "cosine"):
        km = V.KMeans(V.KMeansConfig(20, dist=metric), cent)
        assert (km.find_nearest_batch(rows) == oracle.kmeans_assign(rows, cent, metric)).all()


@pytest.mark.parametrize("metric", ["l2sqr", "cosine"])
def test_lloyd_bit_exact_given_initial_centroids(V, fixtures, oracle, metric):
    """k_means.rs:108-161 from identical initial centroids: centroids and iteration count bit-exact."""
    rows = np.ascontiguousarray(fixtures["base"][:400])
    for sel, k in (((0, 5), 3), ((100, 104), 16), (None, 7)):
        init = oracle.kmeans_pp_init(rows, k, metric, seed=42, sel=sel)
        want, it_want = oracle.kmeans_lloyd(rows, init, metric, 20, 1e-6, sel=sel)
        km = V.KMeans.from_vec_set(rows, V.KMeansConfig(k, 20, 1e-6, metric, sel), init_centroids=init)
        assert km.iterations == it_want
        assert (bits(km.centroids) == bits(want)).all()


def test_lloyd_u8_and_empty_cluster(V, oracle):
    rows = np.array([[0, 0], [2, 0], [0, 2], [2, 2]], np.uint8)
    init = np.array([[1, 1], [200, 200]], np.uint8)
    km = V.KMeans.from_vec_set(rows, V.KMeansConfig(2, 5, 1e-6, "l2sqr"), init_centroids=init)
    assert km.centroids.tolist() == [[1, 1], [200, 200]] and km.iterations == 1
    rng = np.random.default_rng(9)
    rows = rng.integers(0, 256, (500, 6), dtype=np.uint8)
    init = rows[:4].copy()
    want, it_want = oracle.kmeans_lloyd(rows, init, "l2sqr", 10, 1e-6)
    km = V.KMeans.from_vec_set(rows, V.KMeansConfig(4, 10, 1e-6, "l2sqr"), init_centroids=init)
    assert (km.centroids == want).all() and km.iterations == it_want


def test_kmeans_on_real_set_property(V, fixtures):
    """k_means.rs:241-277 through the GPU path with k-means++ on the GPU weights."""
    rows = np.ascontiguousarray(fixtures["base"][:400])
    km = V.KMeans.from_vec_set(rows, V.KMeansConfig(3, 20, 1e-6, "l2sqr", (0, 5)), np.random.default_rng(42))
    assert km.centroids.shape == (3, 5)
    v = np.zeros(960, np.float32)
    v[:5] = km.centroids[1]
    assert km.find_nearest(v) == 1


# ---- PQ ---------------------------------------------------------------------------------------------------
def test_pq_groups_known_answers(V):
    assert V.pq_groups(6, 2) == [(0, 3), (3, 6)]
    assert V.pq_groups(7, 3) == [(0, 3), (3, 5), (5, 7)]
    with pytest.raises(V.VdbError):
        V.pq_groups(3, 5)


@pytest.mark.parametrize("tag,m,n_bits,dimclip", [("pq240", 240, 4, 960), ("pq7", 7, 4, 13), ("pq5b8", 5, 8, 13)])
@pytest.mark.parametrize("metric", ["l2sqr", "cosine"])
def test_pq_golden_bit_exact(V, fixtures, golden, oracle, tag, m, n_bits, dimclip, metric):
    """Codes, lookup tables and ADC distances bit-exact vs the golden vectors; knn_pq by the parity rule."""
    rows = np.ascontiguousarray(fixtures["base"][:200, :dimclip])
    queries = np.ascontiguousarray(fixtures["test"][:5, :dimclip])
    vs = V.DeviceVecSet(rows, metric)
    pq = V.PQTable(vs, V.PQConfig(n_bits, m, metric), golden[f"{tag}_codebooks"])
    assert (pq.encoded_vec_set == golden[f"{tag}_{metric}_codes"]).all()
    lut, qc = pq.create_lookup(queries)
    assert (bits(lut) == bits(golden[f"{tag}_{metric}_lut"])).all()
    assert (bits(qc) == bits(golden[f"{tag}_{metric}_qcache"])).all()
    adc = pq.adc_distances(queries)
    assert (bits(adc) == bits(golden[f"{tag}_{metric}_adc"])).all()
    got = V.FlatIndex(vs).knn_pq_batch(queries, 10, 40, pq)
    want = (golden[f"{tag}_{metric}_knn_ids"], golden[f"{tag}_{metric}_knn_dist"], np.full(5, 10))
    assert_knn_parity(rows, queries, metric, got, want, oracle)


@pytest.mark.parametrize("metric", ["l2sqr", "cosine"])
def test_pq_precise_when_points_fewer_than_centroids(V, oracle, metric):
    """pq_table.rs:324-372 (ADC == exact distance when every point is its own centroid)."""
    rng = np.random.default_rng(42)
    rows = rng.random((5, 8), dtype=np.float32)
    cbs = []
    for lo, hi in V.pq_groups(8, 2):
        c = np.zeros((16, hi - lo), np.float32)
        c[:5] = rows[:, lo:hi]
        c[5:] = 1e3 + np.arange(11)[:, None]
        cbs.append(c.reshape(-1))
    vs = V.DeviceVecSet(rows, metric)
    pq = V.PQTable(vs, V.PQConfig(4, 2, metric), np.concatenate(cbs))
    adc = pq.adc_distances(rows)
    for i in range(5):
        for j in range(5):
            assert abs(adc[i, j] - oracle.distance(rows[i], rows[j], metric)) < 1e-6


@pytest.mark.parametrize("metric", ["l2sqr", "cosine"])
def test_pq_knn_end_to_end_vs_oracle(V, fixtures, oracle, metric):
    """C4 shape on gist_1000: m=240 4-bit, trained on the GPU; then encode/scan/rerank vs the oracle run on the
    SAME codebooks; ef sweep incl. ef < k and ef > n."""
    base, test = fixtures["base"], fixtures["test"][:16]
    vs = V.DeviceVecSet(base, metric)
    pq = V.PQTable.from_vec_set(vs, base, V.PQConfig(4, 240, metric, 300, 5, 1e-6), np.random.default_rng(42))
    codes = oracle.pq_encode(base, pq.codebooks, 240, 4, metric, nthreads=8)
    assert (pq.encoded_vec_set == codes).all()
    idx = V.FlatIndex(vs)
    for k, ef in ((10, 240), (10, 4), (3, 2000)):
        got = idx.knn_pq_batch(test, k, ef, pq)
        want = oracle.flat_knn_pq(base, codes, pq.codebooks, 240, 4, test, k, ef, metric, nthreads=8)
        assert_knn_parity(base, test, metric, got, want, oracle)


@pytest.mark.parametrize("metric", ["l2sqr", "cosine"])
def test_pq_table_test_replica(V, fixtures, oracle, metric):
    """pq_table.rs:374-438: 64 rows x 13 dims, m = ceil(13/3) = 5 (uneven groups), 4 bits, trained on all rows;
    p90 of |adc - exact| / max(exact, 1) over 20 random pairs < 0.2."""
    rows = np.ascontiguousarray(fixtures["base"][:64, :13])
    rng = np.random.default_rng(42)
    vs = V.DeviceVecSet(rows, metric)
    pq = V.PQTable.from_vec_set(vs, rows, V.PQConfig(4, 5, metric, None, 20, 1e-6), rng)
    adc = pq.adc_distances(rows)
    errs = []
    for _ in range(20):
        i0, i1 = int(rng.integers(0, 64)), int(rng.integers(0, 64))
        expected = oracle.distance(rows[i0], rows[i1], metric)
        errs.append(abs(adc[i1, i0] - expected) / max(expected, 1.0))
    errs.sort()
    assert errs[int(np.ceil(20 * 0.9)) - 1] < 0.2


def test_pq_mismatch_errors(V, fixtures):
    base = np.ascontiguousarray(fixtures["base"][:100, :16])
    vs = V.DeviceVecSet(base, "l2sqr")
    with pytest.raises(ValueError):
        V.PQTable(vs, V.PQConfig(5, 4), np.zeros(16 * 16, np.float32))
    pq = V.PQTable(vs, V.PQConfig(4, 4, "l2sqr"), np.ascontiguousarray(base[:16]).reshape(-1))
    vs.push(base[:3])  # the reference drops the PQ table on every write (metadata_vec_table.rs:65)
    with pytest.raises(V.VdbError):
        V.FlatIndex(vs).knn_pq(base[0], 3, 10, pq)


# ---- IVF --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("metric", ["l2sqr", "cosine"])
def test_ivf_golden(V, fixtures, golden, oracle, metric):
    base, test = fixtures["base"], fixtures["test"][:50]
    vs = V.DeviceVecSet(base, metric)
    ivf = V.IVFIndex(vs, base[7:23])
    assert (ivf.assignment == golden[f"assign16_{metric}"]).all()
    off, mem = oracle.ivf_lists(golden[f"assign16_{metric}"], 16)
    for c, lst in enumerate(ivf.clusters):
        assert (lst == mem[int(off[c]):int(off[c + 1])]).all()
    got = ivf.knn_with_ef_batch(test, 10, 4)
    want = (golden[f"ivf_{metric}_ids"], golden[f"ivf_{metric}_dist"], np.full(50, 10))
    assert_knn_parity(base, test, metric, got, want, oracle)


def test_ivf_reference_unit_shape(V, fixtures, oracle):
    """ivf_index.rs:166-235: 1000x12, nlist=7 trained on a 100-row sample, default nprobe 4, k=6 -> same ids as Flat."""
    base = np.ascontiguousarray(fixtures["base"][:, :12])
    ivf = V.IVFIndex.from_vec_set(None, base, "l2sqr", V.IVFConfig(7, 100, 20, 1e-6), np.random.default_rng(42))
    flat = V.FlatIndex(ivf.vec_set)
    a = [p.index for p in ivf.knn(base[200], 6)]
    b = [p.index for p in flat.knn(base[200], 6)]
    assert a == b
    # and against the oracle on the same centroids, several probe counts incl. nprobe > nlist
    off, mem = oracle.ivf_lists(oracle.kmeans_assign(base, ivf.centroids, "l2sqr"), 7)
    q = np.ascontiguousarray(fixtures["test"][:20, :12])
    for nprobe in (1, 2, 4, 7, 50):
        got = ivf.knn_with_ef_batch(q, 6, nprobe)
        want = oracle.ivf_knn(base, ivf.centroids, off, mem, q, 6, nprobe, "l2sqr")
        assert_knn_parity(base, q, "l2sqr", got, want, oracle)
    with pytest.raises(ValueError):
        ivf.knn_with_ef(base[0], 3, 0)


def test_ivf_c3_shape_small(V, oracle):
    """C3 shape scaled down: nlist=128, nprobe sweep 8..24, k=10, 960-d; counts < k when lists are short."""
    rng = np.random.default_rng(1)
    proto = rng.random((40, 960), dtype=np.float32)
    base = (proto[rng.integers(0, 40, 6000)] + 0.05 * rng.standard_normal((6000, 960))).astype(np.float32)
    q = (proto[rng.integers(0, 40, 12)] + 0.05 * rng.standard_normal((12, 960))).astype(np.float32)
    cent = np.ascontiguousarray(base[rng.permutation(6000)[:128]])
    vs = V.DeviceVecSet(base, "l2sqr")
    ivf = V.IVFIndex(vs, cent)
    a = oracle.kmeans_assign(base, cent, "l2sqr", nthreads=8)
    assert (ivf.assignment == a).all()
    off, mem = oracle.ivf_lists(a, 128)
    for nprobe in (1, 8, 16, 24):
        got = ivf.knn_with_ef_batch(q, 10, nprobe)
        want = oracle.ivf_knn(base, cent, off, mem, q, 10, nprobe, "l2sqr", nthreads=8)
        assert_knn_parity(base, q, "l2sqr", got, want, oracle)


@pytest.mark.parametrize("metric", ["l2sqr", "cosine"])
@pytest.mark.parametrize("m", [8, 7])
def test_pq_global_threshold_scan_large_shard(V, oracle, metric, m):
    """Shards >= 65536 rows use the global-threshold ADC scan for batches: same candidates as the reference scan
    (checked through knn_pq vs the oracle on identical codebooks), incl. odd m, many exact ADC ties and ef sweeps."""
    rng = np.random.default_rng(21 + m)
    n, dim = 70_000, 32
    base = rng.random((n, dim), dtype=np.float32)
    base[5000:5300] = base[17]                      # 300 identical rows: ADC ties decided by id
    q = rng.random((9, dim), dtype=np.float32)
    q[0] = base[17]
    books = np.concatenate([np.ascontiguousarray(base[100:116, lo:hi]).reshape(-1) for lo, hi in V.pq_groups(dim, m)])
    vs = V.DeviceVecSet(base, metric)
    pq = V.PQTable(vs, V.PQConfig(4, m, metric), books)
    codes = oracle.pq_encode(base, books, m, 4, metric, nthreads=8)
    assert (pq.encoded_vec_set == codes).all()
    idx = V.FlatIndex(vs)
    for k, ef in ((10, 50), (10, 300), (5, 1200), (3, 1)):
        got = idx.knn_pq_batch(q, k, ef, pq)
        want = oracle.flat_knn_pq(base, codes, books, m, 4, q, k, ef, metric, nthreads=8)
        assert_knn_parity(base, q, metric, got, want, oracle)
    # single query (per-CTA kernel) and the batch (global-threshold kernel) agree exactly
    one = idx.knn_pq_batch(q[:1], 10, 300, pq)
    many = idx.knn_pq_batch(q, 10, 300, pq)
    assert (one[0][0] == many[0][0]).all() and (one[1][0].view(np.uint32) == many[1][0].view(np.uint32)).all()


@pytest.mark.parametrize("metric", ["l2sqr", "cosine"])
def test_ivf_tensor_probe_scan_large_shard(V, oracle, metric):
    """Shards >= 65536 rows scan the probed lists on the tensor cores for batches of >= 16 queries (list-ordered TF32
    rows, gathered query groups, exact rerank): ids and distance bits must equal the FP32 list scan (batches of < 16
    take it) and the oracle, incl. duplicate rows (ties by id), an empty list, k larger than some visit sets, ragged
    dims and short probe counts."""
    rng = np.random.default_rng(77)
    n, dim, nlist = 70_000, 100, 48
    proto = rng.random((nlist - 1, dim), dtype=np.float32)
    base = (proto[rng.integers(0, nlist - 1, n)] + 0.08 * rng.standard_normal((n, dim))).astype(np.float32)
    base[3000:3040] = base[11]                       # ties decided by id
    cent = np.concatenate([proto, np.full((1, dim), 50.0, np.float32)])   # last centroid: nobody's nearest -> empty list
    q = (proto[rng.integers(0, nlist - 1, 40)] + 0.08 * rng.standard_normal((40, dim))).astype(np.float32)
    q[0] = base[11]
    vs = V.DeviceVecSet(base, metric)
    ivf = V.IVFIndex(vs, cent)
    a = oracle.kmeans_assign(base, cent, metric, nthreads=8)
    assert (ivf.assignment == a).all()
    off, mem = oracle.ivf_lists(a, nlist)
    for k, nprobe in ((10, 4), (100, 1), (3, 48), (300, 2)):
        got = ivf.knn_with_ef_batch(q, k, nprobe)
        want = oracle.ivf_knn(base, cent, off, mem, q, k, nprobe, metric, nthreads=8)
        assert_knn_parity(base, q, metric, got, want, oracle)
        fp32 = ivf.knn_with_ef_batch(q[:9], k, nprobe)     # < 16 queries: FP32 list-major scan
        assert (fp32[0] == got[0][:9]).all()
        assert (fp32[1].view(np.uint32) == got[1][:9].view(np.uint32)).all()
        assert (fp32[2] == got[2][:9]).all()


def test_ivf_tensor_probe_scan_u8(V, oracle):
    """u8 rows (exact in TF32) through the tensor probe scan: same ids / distance bits as the FP32 list scan and the oracle."""
    rng = np.random.default_rng(78)
    n, dim, nlist = 66_000, 72, 20
    proto = rng.integers(0, 256, (nlist, dim))
    base = np.clip(proto[rng.integers(0, nlist, n)] + rng.integers(-30, 31, (n, dim)), 0, 255).astype(np.uint8)
    base[100:130] = base[7]
    cent = proto.astype(np.uint8)
    q = np.clip(proto[rng.integers(0, nlist, 33)] + rng.integers(-30, 31, (33, dim)), 0, 255).astype(np.uint8)
    q[1] = base[7]
    vs = V.DeviceVecSet(base, "l2sqr")
    ivf = V.IVFIndex(vs, cent)
    a = oracle.kmeans_assign(base, cent, "l2sqr", nthreads=8)
    assert (ivf.assignment == a).all()
    off, mem = oracle.ivf_lists(a, nlist)
    for k, nprobe in ((10, 3), (50, 1)):
        got = ivf.knn_with_ef_batch(q, k, nprobe)
        want = oracle.ivf_knn(base, cent, off, mem, q, k, nprobe, "l2sqr", nthreads=8)
        assert_knn_parity(base, q, "l2sqr", got, want, oracle)
        fp32 = ivf.knn_with_ef_batch(q[:9], k, nprobe)
        assert (fp32[0] == got[0][:9]).all() and (fp32[1].view(np.uint32) == got[1][:9].view(np.uint32)).all()


@pytest.mark.parametrize("m", [8, 7, 13])
def test_pq_tensor_filter_large_shard(V, oracle, m):
    """Batches of >= 32 queries on shards >= 65536 rows (4-bit, L2Sqr) prune the ADC scan on the tensor cores (bf16
    one-hot contraction, lower bound) and re-evaluate the survivors with the reference arithmetic: candidates, ids and
    distance bits must equal the FP32 scan's / the oracle's, incl. odd m (padded groups), exact ADC ties and a ragged
    last row tile."""
    rng = np.random.default_rng(31 + m)
    n, dim = 70_001, 40
    base = rng.random((n, dim), dtype=np.float32)
    base[5000:5300] = base[17]                      # 300 identical rows: ADC ties decided by id
    q = rng.random((70, dim), dtype=np.float32)
    q[0] = base[17]
    books = np.concatenate([np.ascontiguousarray(base[100:116, lo:hi]).reshape(-1) for lo, hi in V.pq_groups(dim, m)])
    vs = V.DeviceVecSet(base, "l2sqr")
    pq = V.PQTable(vs, V.PQConfig(4, m, "l2sqr"), books)
    codes = oracle.pq_encode(base, books, m, 4, "l2sqr", nthreads=8)
    assert (pq.encoded_vec_set == codes).all()
    idx = V.FlatIndex(vs)
    for k, ef in ((10, 50), (10, 300), (5, 1200), (3, 1)):
        got = idx.knn_pq_batch(q, k, ef, pq)
        want = oracle.flat_knn_pq(base, codes, books, m, 4, q, k, ef, "l2sqr", nthreads=8)
        assert_knn_parity(base, q, "l2sqr", got, want, oracle)
        few = idx.knn_pq_batch(q[:9], k, ef, pq)     # < 32 queries: FP32 global-threshold scan
        assert (few[0] == got[0][:9]).all() and (few[1].view(np.uint32) == got[1][:9].view(np.uint32)).all()


@pytest.mark.parametrize("dim,m,dtype", [(64, 16, np.float32), (960, 240, np.float32), (128, 32, np.uint8)])
def test_pq_decoded_contraction_large_shard(V, oracle, dim, m, dtype, monkeypatch):
    """Sub-vectors of 4 dimensions (m = dim / 4, the reference's bench setting): batches prune the ADC scan with the
    contraction over rows DECODED on the fly (pq_dec.cu). Candidates, ids and distance bits must equal the one-hot
    contraction's, the FP32 scan's and the oracle's, incl. exact ADC ties, a ragged last row tile, a ragged last query
    tile and queries far outside the data."""
    rng = np.random.default_rng(77 + m)
    n = 70_001
    scale = 255 if dtype == np.uint8 else 1
    base = (rng.random((n, dim)) * scale).astype(dtype)
    base[5000:5300] = base[17]                      # 300 identical rows: ADC ties decided by id
    nq = 300 if dim <= 128 else 70
    q = (rng.random((nq, dim)) * scale).astype(dtype)
    q[0] = base[17]
    if dtype == np.float32:
        q[1] *= 40.0                                # a query far away: large norms, loose bounds
        q[2] *= 1e-3
    books = np.concatenate([np.ascontiguousarray(base[100:116, lo:hi]).reshape(-1) for lo, hi in V.pq_groups(dim, m)])
    vs = V.DeviceVecSet(base, "l2sqr")
    pq = V.PQTable(vs, V.PQConfig(4, m, "l2sqr"), books)
    codes = oracle.pq_encode(base, books, m, 4, "l2sqr", nthreads=8)
    assert (pq.encoded_vec_set == codes).all()
    idx = V.FlatIndex(vs)
    for k, ef in ((10, 50), (10, 300), (3, 1)):
        got = idx.knn_pq_batch(q, k, ef, pq)
        monkeypatch.setenv("VDB_PQ_NO_DECODE", "1")
        onehot = idx.knn_pq_batch(q, k, ef, pq)
        monkeypatch.delenv("VDB_PQ_NO_DECODE")
        assert (onehot[0] == got[0]).all() and (onehot[1].view(np.uint32) == got[1].view(np.uint32)).all()
        few = idx.knn_pq_batch(q[:9], k, ef, pq)     # < 32 queries: FP32 global-threshold scan
        assert (few[0] == got[0][:9]).all() and (few[1].view(np.uint32) == got[1][:9].view(np.uint32)).all()
        nchk = 24
        want = oracle.flat_knn_pq(base, codes, books, m, 4, q[:nchk], k, ef, "l2sqr", nthreads=8)
        assert_knn_parity(base, q[:nchk], "l2sqr", tuple(a[:nchk] for a in got), want, oracle)


@pytest.mark.parametrize("metric", ["l2sqr", "cosine"])
@pytest.mark.parametrize("dtype", [np.float32, np.uint8])
def test_batched_pq_training_bit_exact_vs_per_group(V, oracle, metric, dtype):
    """vdb_pq_train_ds trains every group's k-means in one launch. Given the same initial centroids it must equal the
    per-group Lloyd (vdb_kmeans_train_ds) and the oracle bit for bit: uneven groups (pq_groups(43, 9)), empty clusters
    (duplicate initial centroids), u8 rows (Rust `as` cast of the means), early convergence."""
    from lab_1806_vec_db_b200.index import train_codebooks
    rng = np.random.default_rng(55)
    n, dim, m = 1500, 43, 9
    rows = (rng.random((n, dim)) * (255 if dtype == np.uint8 else 1)).astype(dtype)
    rows[100:400] = rows[7]                              # heavy duplicates -> early convergence in some groups
    cfg = V.PQConfig(4, m, metric, None, 12, 1e-6)
    init, parts = [], []
    vs = V.DeviceVecSet(rows, metric)
    for lo, hi in V.pq_groups(dim, m):
        ini = np.ascontiguousarray(rows[rng.integers(0, n, 16), lo:hi])
        ini[3] = ini[2]                                  # duplicate centroid: the higher id stays empty forever
        init.append(ini.reshape(-1))
        km = V.KMeans.from_vec_set(vs, V.KMeansConfig(16, 12, 1e-6, metric, (lo, hi)), rng, ini)
        parts.append(km.centroids.reshape(-1))
        want, _ = oracle.kmeans_lloyd(rows, ini, metric, 12, 1e-6, sel=(lo, hi))
        assert (np.asarray(want).reshape(-1).view(np.uint8) == km.centroids.reshape(-1).view(np.uint8)).all()
    got = train_codebooks(vs, cfg, rng, np.concatenate(init))
    assert (got.view(np.uint8) == np.concatenate(parts).view(np.uint8)).all()


def test_batched_pq_training_end_to_end(V, fixtures, oracle):
    """k-means++ + Lloyd for all groups in one launch: every initial centroid is a training row, quantisation error is
    in line with the per-group path, and PQTable.from_vec_set (which now uses it) passes the reference's own test."""
    from lab_1806_vec_db_b200.index import train_codebooks
    base = fixtures["base"]
    rng = np.random.default_rng(1)
    vs = V.DeviceVecSet(base, "l2sqr")
    cfg = V.PQConfig(4, 240, "l2sqr", None, 20, 1e-6)
    books = train_codebooks(vs, cfg, rng)
    assert books.shape == (240 * 16 * 4,) and np.isfinite(books).all()
    assert (train_codebooks.last_iterations >= 1).all() and (train_codebooks.last_iterations <= 20).all()
    pq = V.PQTable(vs, cfg, books)
    codes = oracle.pq_encode(base, books, 240, 4, "l2sqr", nthreads=8)
    assert (pq.encoded_vec_set == codes).all()
    # quantisation error vs the per-group trainer on the same data
    def qerr(bk):
        c = oracle.pq_encode(base, bk, 240, 4, "l2sqr", nthreads=8)
        rec = np.zeros_like(base)
        o = 0
        for g, (lo, hi) in enumerate(V.pq_groups(960, 240)):
            cb = bk[o:o + 16 * (hi - lo)].reshape(16, hi - lo)
            o += 16 * (hi - lo)
            code = (c[:, g // 2] >> 4) if g % 2 else (c[:, g // 2] & 15)
            rec[:, lo:hi] = cb[code]
        return float(((rec - base) ** 2).sum(1).mean())
    per_group = np.concatenate([V.KMeans.from_vec_set(vs, V.KMeansConfig(16, 20, 1e-6, "l2sqr", (lo, hi)), rng).centroids.reshape(-1)
                                for lo, hi in V.pq_groups(960, 240)])
    assert qerr(books) <= 1.05 * qerr(per_group)
    # 0 iterations allowed: only k-means++ -> every centroid is a training row's sub-vector
    cfg0 = V.PQConfig(4, 8, "l2sqr", None, 0, 1e-6)
    b0 = train_codebooks(vs, cfg0, np.random.default_rng(2))
    o = 0
    for lo, hi in V.pq_groups(960, 8):
        cb = b0[o:o + 16 * (hi - lo)].reshape(16, hi - lo)
        o += 16 * (hi - lo)
        for c in cb:
            assert (np.abs(base[:, lo:hi] - c).max(1) == 0).any()


def test_pq_tensor_filter_m320(V, oracle):
    """m = 320 (the reference's published HNSW+PQ / Flat+PQ setting, config/bench_10000_pq_flat.toml): 160-byte codes,
    80 k-blocks - the code tile is sized from the actual encoded_dim."""
    rng = np.random.default_rng(320)
    n, dim, m = 66_000, 330, 320
    base = rng.random((n, dim), dtype=np.float32)
    q = rng.random((40, dim), dtype=np.float32)
    books = np.concatenate([np.ascontiguousarray(base[200:216, lo:hi]).reshape(-1) for lo, hi in V.pq_groups(dim, m)])
    vs = V.DeviceVecSet(base, "l2sqr")
    pq = V.PQTable(vs, V.PQConfig(4, m, "l2sqr"), books)
    codes = oracle.pq_encode(base, books, m, 4, "l2sqr", nthreads=8)
    assert pq.encoded_vec_set.shape == (n, 160) and (pq.encoded_vec_set == codes).all()
    idx = V.FlatIndex(vs)
    for k, ef in ((10, 100), (10, 200)):
        got = idx.knn_pq_batch(q, k, ef, pq)
        want = oracle.flat_knn_pq(base, codes, books, m, 4, q, k, ef, "l2sqr", nthreads=8)
        assert_knn_parity(base, q, "l2sqr", got, want, oracle)


# ---- the reference's random stream (rand_compat.StdRng) driving the training entry points ------------------------------
def test_kmeans_pp_under_the_restated_reference_stream(V):
    """k-means++ (k_means.rs:61-87) with the reference's draw sequence: the GPU's weight updates (`vdb_kmeans_pp_weights`)
    must be the bits of the sequential-f32 emulation, so the same stream picks the same rows; the training entry points
    accept the stream in place of a numpy Generator (random_sample = the restated Fisher-Yates shuffle)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import np_emul as E
    from lab_1806_vec_db_b200.index import k_means_init, sample_indices
    from lab_1806_vec_db_b200.rand_compat import StdRng, k_means_init_indices
    rng = np.random.default_rng(5)
    rows = rng.random((600, 40), dtype=np.float32)
    k = 12
    for sel in (None, (8, 24)):
        lo, hi = sel if sel else (0, rows.shape[1])
        w = np.full(len(rows), np.inf, np.float32)

        def update(idx):
            np.minimum(w, E.l2sqr(np.broadcast_to(rows[idx, lo:hi], (len(rows), hi - lo)), rows[:, lo:hi]), out=w)
            return w

        want = k_means_init_indices(update, len(rows), k, StdRng.seed_from_u64(42))
        cfg = V.KMeansConfig(k, 5, 1e-6, "l2sqr", sel)
        got = k_means_init(rows, cfg, StdRng.seed_from_u64(42))
        assert (bits(got) == bits(rows[want, lo:hi])).all()
        km = V.KMeans.from_vec_set(rows, cfg, StdRng.seed_from_u64(42))
        assert km.centroids.shape == (k, hi - lo)
    # random_sample: the first entries of the shuffled index vector
    assert (sample_indices(StdRng.seed_from_u64(1), 600, 100) == StdRng.seed_from_u64(1).shuffle(600)[:100]).all()
    vs = V.DeviceVecSet(rows, "l2sqr")
    pq = V.PQTable.from_vec_set(vs, rows, V.PQConfig(4, 10, "l2sqr", 200, 5, 1e-6), StdRng.seed_from_u64(42))
    pq2 = V.PQTable.from_vec_set(vs, rows, V.PQConfig(4, 10, "l2sqr", 200, 5, 1e-6), StdRng.seed_from_u64(42))
    assert (bits(pq.codebooks) == bits(pq2.codebooks)).all() and (pq.encoded_vec_set == pq2.encoded_vec_set).all()
    ivf = V.IVFIndex.from_vec_set(vs, rows, "l2sqr", V.IVFConfig(8, 300, 5, 1e-6), StdRng.seed_from_u64(42))
    assert ivf.centroids.shape == (8, rows.shape[1])
