"""VecDB / MetadataVecTable surface on the GPU backend: the reference's own API smoke tests."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_examples_test_pyo3_script():
    """examples/test_pyo3.py:1-37, statement for statement."""
    from lab_1806_vec_db_b200.table import VecDB
    db = VecDB("./tmp/vec_db")
    for key in db.get_all_keys():
        db.delete_table(key)
    assert len(db.get_all_keys()) == 0
    db.create_table_if_not_exists("table_1", 4)
    db.add("table_1", [1.0, 0.0, 0.0, 0.0], {"content": "a"})
    db.add("table_1", [0.0, 1.0, 0.0, 0.0], {"content": "b"})
    db.build_hnsw_index("table_1")
    db.add("table_1", [0.0, 0.0, 1.0, 0.0], {"content": "c"})
    db.add("table_1", [0.0, 0.0, 1.0, 1.0], {"content": "d", "type": "oops"})
    assert db.has_hnsw_index("table_1"), "Add operation should not clear HNSW index"
    db.delete("table_1", {"type": "oops"})
    assert db.get_len("table_1") == 3
    assert not db.has_hnsw_index("table_1"), "HNSW index should be cleared when a vector is deleted"
    db.build_hnsw_index("table_1")
    db.build_pq_table("table_1")
    result = db.search("table_1", [1.0, 0.0, 0.0, 0.0], 3, None, 0.5)
    assert len(result) == 1 and result[0][0]["content"] == "a"


def test_database_mod_rs_scenario():
    """database/mod.rs:543-610: PQ + HNSW search with upper_bound = 0.5 returns exactly ["c"]."""
    from lab_1806_vec_db_b200.table import VecDB
    db = VecDB()
    assert db.create_table_if_not_exists("t", 4, "cosine") and not db.create_table_if_not_exists("t", 4, "cosine")
    db.batch_add("t", [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]], [{"content": c} for c in "abc"])
    assert db.get_len("t") == 3 and db.get_dim("t") == 4 and db.get_dist("t") == "cosine"
    db.build_pq_table("t")
    assert db.has_pq_table("t")
    db.add("t", [0, 0, 0, 1], {"content": "d"})
    assert not db.has_pq_table("t")            # the PQ table is dropped on every write
    db.build_pq_table("t")
    db.build_hnsw_index("t")
    res = db.search("t", [0.0, 0.0, 1.0, 0.0], 3, 10, 0.5)
    assert [m["content"] for m, _ in res] == ["c"] and abs(res[0][1]) < 1e-6
    db.delete("t", {"content": "a"})           # swap_remove: "d" takes slot 0
    assert not db.has_pq_table("t") and not db.has_hnsw_index("t")
    assert [m["content"] for _, m in db.extract_data("t")] == ["d", "b", "c"]
    res = db.search("t", [0, 0, 0, 1.0], 1)
    assert res[0][0]["content"] == "d"
    with pytest.raises(ValueError):
        db.create_table_if_not_exists("x", 4, "manhattan")
    with pytest.raises(RuntimeError):
        db.search("missing", [0, 0, 0, 0], 1)


def test_search_variants_against_oracle(fixtures, oracle):
    from lab_1806_vec_db_b200.table import MetadataVecTable
    base = fixtures["base"][:300]
    t = MetadataVecTable(960, "l2sqr", np.random.default_rng(42))
    t.batch_add(base, [{"i": str(i)} for i in range(300)])
    q = fixtures["test"][0]
    want = oracle.flat_knn(base, q.reshape(1, -1), 5, "l2sqr")
    got = t.search(q, 5)
    assert [int(m["i"]) for m, _ in got] == want[0][0].tolist()
    assert [int(m["i"]) for m, _ in t.search(q, 5, ef=50)] == want[0][0].tolist()   # no PQ: ef ignored
    ub = float(want[1][0][2]) * (1 + 1e-5)  # the GPU distance may differ from the oracle by an ulp
    assert len(t.search(q, 5, upper_bound=ub)) == 3                                   # distance <= upper_bound
    t.build_pq_table(0.5, 8, 240)                                                     # n_bits is forced to 4
    assert t.pq_table.config.n_bits == 4 and t.pq_table.config.k_means_size == 150
    got = t.search(q, 5, ef=300)                                                      # ef >= n: rerank of everything
    assert [int(m["i"]) for m, _ in got] == want[0][0].tolist()


def test_hnsw_variants_of_the_search_policy(fixtures, oracle):
    """DynamicIndex::HNSW arms (dynamic_index.rs:63-93): knn -> default ef, knn_with_ef, knn_pq on the graph; an add
    after the build is visible to the next search; delete clears the graph."""
    from lab_1806_vec_db_b200.table import MetadataVecTable
    base = fixtures["base"][:400]
    t = MetadataVecTable(960, "l2sqr", np.random.default_rng(7))
    t.batch_add(base[:350], [{"i": str(i)} for i in range(350)])
    t.build_hnsw_index(64)
    assert t.has_hnsw_index()
    q = fixtures["test"][1]
    want = oracle.flat_knn(base[:350], q.reshape(1, -1), 5, "l2sqr")[0][0].tolist()
    assert [int(m["i"]) for m, _ in t.search(q, 5)] == want                 # default ef = ef_construction / 2
    assert [int(m["i"]) for m, _ in t.search(q, 5, ef=200)] == want
    t.build_pq_table(0.5, 4, 240)
    got = [int(m["i"]) for m, _ in t.search(q, 5, ef=350)]                  # graph walk with ADC, exact resort
    assert len(set(got) & set(want)) >= 4
    t.batch_add(base[350:], [{"i": str(i)} for i in range(350, 400)])      # drops PQ, graph follows the new rows
    assert t.has_hnsw_index() and not t.has_pq_table()
    got = t.search(base[399], 1, ef=100)
    assert got[0][0]["i"] == "399" and abs(got[0][1]) < 1e-5
    t.delete({"i": "0"})
    assert not t.has_hnsw_index()
