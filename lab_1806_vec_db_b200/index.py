"""Host-side mirror of the reference's index_algorithm surface for the GPU path.

Names follow the reference: FlatIndex.from_vec_set / knn (src/index_algorithm/flat_index.rs:48-70),
CandidatePair (src/index_algorithm/candidate_pair.rs:10-40). `knn_batch` is the additive batch entry
(the reference is single-query; examples/bench.rs:414-418 batches with rayon instead).
"""
import ctypes as C
from typing import List, NamedTuple

import numpy as np

from . import _lib as L


class CandidatePair(NamedTuple):
    """(index, distance) as returned by IndexKNN::knn; ordered by (distance, index)."""
    index: int
    distance: float


def _as_rows(a, dtype=None):
    a = np.asarray(a)
    if dtype is not None and a.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {a.dtype}")
    if a.ndim == 1:
        a = a.reshape(1, -1)
    return np.ascontiguousarray(a)


class DeviceVecSet:
    """Device mirror of VecSet<T> (reference src/vec_set.rs:15-30) — one row shard in HBM."""

    def __init__(self, rows, dist="l2sqr", id_base=0):
        rows = np.ascontiguousarray(rows)
        if rows.ndim != 2:
            raise ValueError("rows must be [n, dim]")
        self.dtype = rows.dtype
        self.dim = int(rows.shape[1])
        self.metric = L.metric_code(dist)
        self.id_base = int(id_base)
        self._h = C.c_void_p()
        L.check(L.lib().vdb_dataset_create(L.ptr(rows), rows.shape[0], self.dim, L.dtype_code(rows),
                                           self.metric, self.id_base, C.byref(self._h)))
        self._keepalive = None

    @classmethod
    def from_device(cls, d_ptr, n, dim, pitch, dtype, dist="l2sqr", id_base=0, keepalive=None):
        """Adopts rows already resident in HBM (e.g. a torch tensor's data_ptr())."""
        self = cls.__new__(cls)
        self.dtype = np.dtype(dtype)
        self.dim = int(dim)
        self.metric = L.metric_code(dist)
        self.id_base = int(id_base)
        self._h = C.c_void_p()
        code = L.F32 if self.dtype == np.float32 else L.U8
        L.check(L.lib().vdb_dataset_create_dev(C.c_void_p(int(d_ptr)), n, dim, pitch, code, self.metric,
                                               self.id_base, C.byref(self._h)))
        self._keepalive = keepalive
        return self

    def __len__(self):
        n = C.c_uint64()
        L.check(L.lib().vdb_dataset_len(self._h, C.byref(n)))
        return int(n.value)

    def push(self, rows):
        """VecSet::push (src/vec_set.rs:113-118), batched."""
        rows = _as_rows(rows, self.dtype)
        if rows.shape[1] != self.dim:
            raise ValueError("The dimension of the vector doesn't match.")
        L.check(L.lib().vdb_dataset_append(self._h, L.ptr(rows), rows.shape[0]))

    def swap_remove(self, idx):
        """VecSet::swap_remove (src/vec_set.rs:131-137)."""
        L.check(L.lib().vdb_dataset_swap_remove(self._h, int(idx)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            L.lib().vdb_dataset_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _pairs(ids, dist, counts):
    out = []
    for q in range(ids.shape[0]):
        c = int(counts[q])
        out.append([CandidatePair(int(ids[q, j]), float(dist[q, j])) for j in range(c)])
    return out


class FlatIndex:
    """FlatIndex<T> (reference src/index_algorithm/flat_index.rs:17-57) backed by the GPU scan."""

    def __init__(self, vec_set: DeviceVecSet):
        self.vec_set = vec_set

    @classmethod
    def from_vec_set(cls, vec_set, dist="l2sqr", config=None, rng=None):
        """IndexFromVecSet::from_vec_set (flat_index.rs:59-70); `config`/`rng` are unused as there."""
        if not isinstance(vec_set, DeviceVecSet):
            vec_set = DeviceVecSet(vec_set, dist)
        return cls(vec_set)

    def __len__(self):
        return len(self.vec_set)

    def dim(self):
        return self.vec_set.dim

    def knn_batch(self, queries, k):
        """Returns (ids [nq,k] u64, dist [nq,k] f32, counts [nq] u32)."""
        vs = self.vec_set
        q = _as_rows(queries, vs.dtype)
        if q.shape[1] != vs.dim:
            raise ValueError("The dimension of the query doesn't match.")
        nq = q.shape[0]
        ids = np.full((nq, k), np.iinfo(np.uint64).max, np.uint64)
        dist = np.full((nq, k), np.nan, np.float32)
        counts = np.zeros(nq, np.uint32)
        L.check(L.lib().vdb_flat_knn(vs._h, L.ptr(q), nq, k, L.ptr(ids), L.ptr(dist), L.ptr(counts)))
        return ids, dist, counts

    def knn(self, query, k) -> List[CandidatePair]:
        """IndexKNN::knn (flat_index.rs:48-57)."""
        ids, dist, counts = self.knn_batch(np.asarray(query).reshape(1, -1), k)
        return _pairs(ids, dist, counts)[0]


def calc_dist(a, b, dist="cosine"):
    """calc_dist(a, b, dist="cosine") of the pyo3 layer (reference src/pyo3/mod.rs:43-48)."""
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
    b = np.ascontiguousarray(b, dtype=np.float32).reshape(-1)
    if a.size != b.size:
        raise ValueError("The dimension of the vectors doesn't match.")
    out = np.zeros(1, np.float32)
    L.check(L.lib().vdb_calc_dist(L.ptr(a), L.ptr(b), 1, a.size, L.F32, L.metric_code(dist), L.ptr(out)))
    return float(out[0])
