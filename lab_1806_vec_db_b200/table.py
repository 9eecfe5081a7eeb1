"""MetadataVecTable search policy + mutation hooks and the VecDB call surface on top of the GPU backend
(SURVEY.md section 8f rank 3). Mirrors reference src/database/metadata_vec_table.rs:63-212 and the pyo3 method
names/arguments (src/pyo3/mod.rs, lab_1806_vec_db.pyi) so `VecDB.search / build_pq_table / build_hnsw_index` callers
can switch without code changes. Storage, locking and autosave are out of scope: tables live in memory.

`build_hnsw_index` builds the graph on the device (hnsw.cu), the search policy follows DynamicIndex
(src/database/dynamic_index.rs:63-93) and vectors added after the build are inserted into the existing graph
(DynamicIndex::HNSW add / batch_add, dynamic_index.rs:44-55) with the same batched pipeline as the build.
"""
import numpy as np

from .index import DeviceVecSet, FlatIndex, HNSWConfig, HNSWIndex, PQConfig, PQTable
from . import _lib as L


def _match(metadata, pattern):
    return all(metadata.get(k) == v for k, v in pattern.items())


class MetadataVecTable:
    def __init__(self, dim, dist="cosine", rng=None):
        L.metric_code(dist)  # ValueError on an invalid distance function (pyo3/mod.rs:15-22)
        self.dim_, self.dist_ = int(dim), dist.lower()
        self.rows = np.zeros((0, self.dim_), np.float32)      # host rows stay authoritative
        self.metadata = []
        self.vec_set = None                                   # GPU mirror, created on first use
        self.pq_table = None
        self._hnsw = None                                     # HNSWIndex, or None (DynamicIndex::Flat)
        self._hnsw_efc = None
        self.rng = rng if rng is not None else np.random.default_rng()  # from_entropy, metadata_vec_table.rs:18

    def __len__(self):
        return len(self.metadata)

    def dim(self):
        return self.dim_

    def dist(self):
        return self.dist_

    # ---- mutation hooks: the GPU mirror follows push / swap_remove, PQ is dropped on every write ----
    def _mirror(self):
        if self.vec_set is None:
            self.vec_set = DeviceVecSet(self.rows, self.dist_)
        return self.vec_set

    def add(self, vec, metadata):
        self.batch_add([vec], [metadata])

    def batch_add(self, vec_list, metadata_list):
        assert len(vec_list) == len(metadata_list)
        if not len(vec_list):
            return
        new = np.ascontiguousarray(vec_list, np.float32).reshape(len(vec_list), -1)
        if new.shape[1] != self.dim_:
            raise ValueError("The dimension of the vector doesn't match.")
        self.clear_pq_table()                                  # metadata_vec_table.rs:65, 77
        self.metadata.extend(dict(m) for m in metadata_list)
        self.rows = np.concatenate([self.rows, new])
        if self.vec_set is not None:
            self.vec_set.push(new)                             # VecSet::push on the mirror
            if self._hnsw is not None:
                self._hnsw.batch_add_pushed(self.rng)          # DynamicIndex::HNSW batch_add (dynamic_index.rs:50-55)

    def delete(self, pattern):
        self.clear_hnsw_index()                                # :170
        self.clear_pq_table()                                  # :171
        matches = [i for i, m in enumerate(self.metadata) if _match(m, pattern)]
        for i in reversed(matches):                            # swap_remove from the back (:180-185)
            last = len(self.metadata) - 1
            self.metadata[i] = self.metadata[last]
            self.metadata.pop()
            self.rows[i] = self.rows[last]
            self.rows = self.rows[:last]
            if self.vec_set is not None:
                self.vec_set.swap_remove(i)
        return len(matches)

    # ---- indexes ----
    def build_hnsw_index(self, ef_construction=None):
        """metadata_vec_table.rs:84-98: skipped when already built; HNSWConfig defaults (M = 16, ef_construction = 200)."""
        if self._hnsw is not None:
            return
        self._hnsw_efc = ef_construction
        self._build_hnsw()

    def _build_hnsw(self):
        cfg = HNSWConfig(len(self), 200 if self._hnsw_efc is None else self._hnsw_efc, 16)
        if self._hnsw is not None:
            self._hnsw.close()
        self._hnsw = HNSWIndex(self._mirror(), cfg, self.rng)

    def clear_hnsw_index(self):
        if self._hnsw is not None:
            self._hnsw.close()
        self._hnsw = None

    def has_hnsw_index(self):
        return self._hnsw is not None

    def build_pq_table(self, train_proportion=None, n_bits=None, m=None):
        """metadata_vec_table.rs:112-152 (note: n_bits is validated, then 4 is always used, :140)."""
        if self.pq_table is not None:
            return
        if len(self) == 0:
            raise RuntimeError("Cannot build PQ table for an empty table")
        proportion = 0.1 if train_proportion is None else train_proportion
        if proportion <= 0.0 or proportion >= 1.0:
            raise RuntimeError("Train proportion must be in (0, 1)")
        train_size = int(max(np.float32(len(self)) * np.float32(proportion), 1.0))
        n_bits = 4 if n_bits is None else n_bits
        if n_bits not in (4, 8):
            raise RuntimeError("n_bits must be 4 or 8")
        m = -(-self.dim_ // 3) if m is None else m
        if m == 0 or m > self.dim_:
            raise RuntimeError("m must be in 1..=dim")
        cfg = PQConfig(4, m, self.dist_, train_size, 20, 1e-6)
        self.pq_table = PQTable.from_vec_set(self._mirror(), self.rows, cfg, self.rng)

    def clear_pq_table(self):
        if self.pq_table is not None:
            self.pq_table.close()
        self.pq_table = None

    def has_pq_table(self):
        return self.pq_table is not None

    # ---- search policy (metadata_vec_table.rs:194-212) ----
    def search(self, query, k, ef=None, upper_bound=None):
        if len(self) == 0:
            return []
        q = np.ascontiguousarray(query, np.float32).reshape(-1)
        if q.size != self.dim_:
            raise ValueError("The dimension of the query doesn't match.")
        if self._hnsw is not None:
            inner = self._hnsw
        else:
            inner = FlatIndex(self._mirror())
        if ef is not None and self.pq_table is not None:
            results = inner.knn_pq(q, k, ef, self.pq_table)
        elif ef is not None and self._hnsw is not None:
            results = inner.knn_with_ef(q, k, ef)
        else:
            results = inner.knn(q, k)  # Flat ignores ef (dynamic_index.rs:77); HNSW uses its default ef
        ub = np.inf if upper_bound is None else upper_bound
        return [(dict(self.metadata[p.index]), p.distance) for p in results if p.distance <= ub]

    def extract_data(self):
        return [(r.tolist(), dict(m)) for r, m in zip(self.rows, self.metadata)]


class VecDB:
    """In-memory stand-in for the pyo3 `VecDB` class with the same method names and arguments (lab_1806_vec_db.pyi);
    `dir` is accepted and ignored (persistence is out of scope)."""

    def __init__(self, dir=None):
        self.dir = dir
        self.tables = {}

    def create_table_if_not_exists(self, key, dim, dist="cosine"):
        if key in self.tables:
            return False
        self.tables[key] = MetadataVecTable(dim, dist)
        return True

    def _t(self, key):
        if key not in self.tables:
            raise RuntimeError(f"Table {key} not found")
        return self.tables[key]

    def get_len(self, key): return len(self._t(key))
    def get_dim(self, key): return self._t(key).dim()
    def get_dist(self, key): return self._t(key).dist()
    def delete_table(self, key): return self.tables.pop(key, None) is not None
    def get_all_keys(self): return sorted(self.tables)
    def contains_key(self, key): return key in self.tables
    def get_cached_tables(self): return sorted(self.tables)
    def contains_cached(self, key): return key in self.tables
    def remove_cached_table(self, key): pass
    def add(self, key, vec, metadata): self._t(key).add(vec, metadata)
    def batch_add(self, key, vec_list, metadata_list): self._t(key).batch_add(vec_list, metadata_list)
    def delete(self, key, pattern): self._t(key).delete(pattern)
    def search(self, key, query, k, ef=None, upper_bound=None): return self._t(key).search(query, k, ef, upper_bound)
    def extract_data(self, key): return self._t(key).extract_data()
    def build_hnsw_index(self, key, ef_construction=None): self._t(key).build_hnsw_index(ef_construction)
    def clear_hnsw_index(self, key): self._t(key).clear_hnsw_index()
    def has_hnsw_index(self, key): return self._t(key).has_hnsw_index()

    def build_pq_table(self, key, train_proportion=None, n_bits=None, m=None):
        self._t(key).build_pq_table(train_proportion, n_bits, m)

    def clear_pq_table(self, key): self._t(key).clear_pq_table()
    def has_pq_table(self, key): return self._t(key).has_pq_table()
    def force_save(self): pass
