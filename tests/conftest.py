import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

RTOL = 1e-5      # north star: distances within 1e-5 relative, ids exact except ties within 1e-5
ATOL = 1e-6      # absolute floor near zero (the reference's own tests use abs 1e-6, distance/mod.rs:136)
# Cosine distance is 1 - cos: the subtraction cancels, so the 1e-5 relative bar is applied to the O(1) cosine term,
# i.e. as an ABSOLUTE 1e-5 ("for cosine use an absolute floor"). The reference's own
# sequential-f32 cosine is only accurate to ~3e-6 absolute on 960-d data, so no summation order can do better.
ATOL_COSINE = 1e-5


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box)")


def _decode(u16):
    return (u16.astype(np.float64) / 1e4).astype(np.float32)


@pytest.fixture(scope="session")
def fixtures():
    """The reference's data/gist_1000.bin and data/gist_test.bin (losslessly re-encoded)."""
    fx = np.load(os.path.join(GOLDEN, "fixtures.npz"))
    return {"base": _decode(fx["base_u16"]), "test": _decode(fx["test_u16"]),
            "base_sha256": str(fx["base_sha256"]), "test_sha256": str(fx["test_sha256"])}


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "golden.npz"))


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.lib()
    return O


def close(a, b, rtol=RTOL, atol=ATOL):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) <= atol + rtol * np.abs(b)


def assert_knn_parity(base, queries, metric, got, want, oracle, rtol=RTOL, atol=ATOL, rows_of=None):
    """got/want = (ids [nq,k], dist [nq,k], counts [nq]). Parity rule of BASELINE.json:
    counts equal; distances within rtol; ids identical except where the oracle's own distance of
    the returned id ties the oracle's distance at that rank within rtol. Returns the exact-id rate."""
    if metric in ("cosine", 1) and atol == ATOL:
        atol = ATOL_COSINE
    gi, gd, gc = got
    wi, wd, wc = want
    assert gi.shape == wi.shape and gd.shape == wd.shape
    assert (np.asarray(gc) == np.asarray(wc)).all(), "result counts differ"
    exact = 0
    total = 0
    for q in range(gi.shape[0]):
        c = int(wc[q])
        total += c
        assert close(gd[q, :c], wd[q, :c], rtol, atol).all(), (
            f"query {q}: distances differ: {gd[q, :c]} vs {wd[q, :c]}")
        assert len(set(gi[q, :c].tolist())) == c, f"query {q}: duplicate ids"
        same = gi[q, :c].astype(np.int64) == wi[q, :c].astype(np.int64)
        exact += int(same.sum())
        for j in np.nonzero(~same)[0]:
            row = base[int(gi[q, j])] if rows_of is None else rows_of(int(gi[q, j]))
            d = oracle.distance(queries[q], row, metric)
            assert close(d, wd[q, j], rtol, atol), (
                f"query {q} rank {j}: id {gi[q, j]} (oracle distance {d}) is not a tie of "
                f"id {wi[q, j]} (distance {wd[q, j]})")
    return exact / max(total, 1)


def have_gpu():
    """True when a CUDA device is visible. A missing or broken libvdb_b200.so (load failure, symbol mismatch) is NOT
    "no GPU": it propagates, so a broken build fails the suite instead of skipping it."""
    import ctypes as C
    from lab_1806_vec_db_b200 import _lib as L
    n = C.c_int(0)
    rc = L.lib().vdb_device_count(C.byref(n))
    return rc == 0 and n.value > 0
