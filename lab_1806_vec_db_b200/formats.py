"""Data formats either side of the hot path.

Everything here is host-side byte shuffling so that indexes built by this backend can be loaded by the unmodified
Rust crate and vice versa. The layouts follow bincode 1.3.3's default configuration as used by the reference
(`bincode::serialize_into`): little-endian, fixed-width integers, `usize` as u64, enum variant tag as u32,
`Option` tag as u8, `Vec<T>` / `String` as u64 length + items, structs as their fields in declaration order.

  raw vectors      headerless little-endian rows                     src/scalar.rs:73-108
  fvecs            per vector: i32 dim + dim f32                      src/bin/convert_fvecs.rs:29-48
  GroundTruth      Vec<GroundTruthRow{knn_indices: Vec<usize>}>       src/index_algorithm/candidate_pair.rs:111-191
  VecSet<T>        {dim: usize, data: Vec<T>}                         src/vec_set.rs:15-20
  KMeans<T>        {config: KMeansConfig, centroids: VecSet<T>}       src/distance/k_means.rs:15-37
  PQTable<T>       pq_table.rs:116-137
  IVFIndex<T>      ivf_index.rs:34-47 (vec_set emptied by save_without_vec_set, :109-121)
  Flat index file  the 4-byte DistanceAlgorithm tag                   flat_index.rs:72-83
  ResultList       TOML written by examples/bench.rs:312-368

The reference ships no index files, so these layouts are pinned by bincode's specification and by byte-level
known-answer tests (tests/test_formats_cpu.py), not by reference artefacts.
"""
import struct
from typing import List, NamedTuple, Optional

import numpy as np

DIST_TAG = {"l2sqr": 0, "cosine": 1}          # DistanceAlgorithm variant order, src/distance/mod.rs:18-28
DIST_NAME = {0: "l2sqr", 1: "cosine"}
DIST_TOML = {"L2Sqr": "l2sqr", "Cosine": "cosine"}


# ---- raw vectors / fvecs --------------------------------------------------------------------------------------
def load_raw(path, dim, dtype=np.float32, limit=None):
    """BinaryScalar::from_binary_file (scalar.rs:89-98) + VecDataConfig.limit (config.rs:31-52)."""
    # the reference reads min(limit * dim, file scalars) SCALARS (file_size_limit, scalar.rs:78-82) and only then checks
    # the row multiple (vec_set.rs:35-38): a trailing partial row beyond the limit is never looked at
    count = -1 if limit is None else int(limit) * int(dim)
    a = np.fromfile(path, dtype=np.dtype(dtype).newbyteorder("<"), count=count)
    if a.size % dim:
        raise ValueError("The length of the data is not a multiple of the dimension.")
    return np.ascontiguousarray(a.reshape(-1, dim), dtype=dtype)


def save_raw(path, rows):
    np.ascontiguousarray(rows).astype(np.asarray(rows).dtype.newbyteorder("<"), copy=False).tofile(path)


def read_fvecs(path, limit=None):
    """convert_fvecs.rs:29-48: each vector is an i32 dimension followed by that many f32."""
    raw = np.fromfile(path, dtype="<i4")
    if raw.size == 0:
        return np.zeros((0, 0), np.float32)
    dim = int(raw[0])
    rec = raw.reshape(-1, dim + 1)
    if (rec[:, 0] != dim).any():
        raise ValueError("inconsistent dimensions in fvecs file")
    rows = rec[:, 1:].view("<f4")
    return np.ascontiguousarray(rows[:limit] if limit is not None else rows, np.float32)


# ---- bincode primitives ------------------------------------------------------------------------------------------
class _W:
    def __init__(self):
        self.b = bytearray()

    def u8(self, v): self.b += struct.pack("<B", v)
    def u32(self, v): self.b += struct.pack("<I", v)
    def u64(self, v): self.b += struct.pack("<Q", v)
    def f32(self, v): self.b += struct.pack("<f", v)

    def opt_u64(self, v):
        if v is None:
            self.u8(0)
        else:
            self.u8(1)
            self.u64(v)

    def array(self, a, dtype):
        a = np.ascontiguousarray(a, dtype=np.dtype(dtype).newbyteorder("<")).reshape(-1)
        self.u64(a.size)
        self.b += a.tobytes()


class _R:
    def __init__(self, data):
        self.d, self.o = memoryview(data), 0

    def _take(self, fmt, n):
        v = struct.unpack_from(fmt, self.d, self.o)[0]
        self.o += n
        return v

    def u8(self): return self._take("<B", 1)
    def u32(self): return self._take("<I", 4)
    def u64(self): return self._take("<Q", 8)
    def f32(self): return self._take("<f", 4)

    def opt_u64(self):
        return self.u64() if self.u8() else None

    def array(self, dtype):
        n = self.u64()
        dt = np.dtype(dtype).newbyteorder("<")
        a = np.frombuffer(self.d, dtype=dt, count=n, offset=self.o).astype(dtype)
        self.o += n * dt.itemsize
        return a

    def done(self):
        if self.o != len(self.d):
            raise ValueError(f"{len(self.d) - self.o} trailing bytes")


# ---- GroundTruth ---------------------------------------------------------------------------------------------------
def dump_ground_truth(rows) -> bytes:
    """rows: iterable of id sequences (the k ids per test query, gen_gnd.rs:54-68)."""
    w = _W()
    rows = list(rows)
    w.u64(len(rows))
    for r in rows:
        w.array(np.asarray(r, np.uint64), np.uint64)
    return bytes(w.b)


def load_ground_truth(data) -> List[np.ndarray]:
    r = _R(data)
    out = [r.array(np.uint64) for _ in range(r.u64())]
    r.done()
    return out


def recall(gnd_row, result_ids) -> float:
    """GroundTruthRow::recall (candidate_pair.rs:127-140)."""
    pred = set(int(i) for i in result_ids)
    return sum(1 for i in gnd_row if int(i) in pred) / len(gnd_row)


# ---- VecSet / KMeans / PQTable / IVFIndex ---------------------------------------------------------------------------
def _w_vecset(w, rows, dtype):
    rows = np.asarray(rows, dtype)
    w.u64(rows.shape[1] if rows.ndim == 2 else 0)
    w.array(rows, dtype)


def _r_vecset(r, dtype):
    dim = r.u64()
    data = r.array(dtype)
    return data.reshape(-1, dim) if dim else data.reshape(0, 0)


class KMeansRecord(NamedTuple):
    k: int
    max_iter: int
    tol: float
    dist: str
    selected: Optional[tuple]
    centroids: np.ndarray


def _w_kmeans(w, km: KMeansRecord, dtype):
    w.u64(km.k)
    w.u64(km.max_iter)
    w.f32(km.tol)
    w.u32(DIST_TAG[km.dist])
    if km.selected is None:
        w.u8(0)
    else:
        w.u8(1)
        w.u64(km.selected[0])   # Range<usize> serialises as {start, end}
        w.u64(km.selected[1])
    _w_vecset(w, km.centroids, dtype)


def _r_kmeans(r, dtype) -> KMeansRecord:
    k, max_iter, tol, dist = r.u64(), r.u64(), r.f32(), DIST_NAME[r.u32()]
    sel = (r.u64(), r.u64()) if r.u8() else None
    return KMeansRecord(k, max_iter, tol, dist, sel, _r_vecset(r, dtype))


class PQTableRecord(NamedTuple):
    n_bits: int
    m: int
    dist: str
    k_means_size: Optional[int]
    k_means_max_iter: int
    k_means_tol: float
    dim: int
    encoded_vec_set: np.ndarray          # [n, encoded_dim] u8
    group_k_means: List[KMeansRecord]
    dist_cache: np.ndarray               # [m * 2**n_bits] f32


def dump_pq_table(t: PQTableRecord, dtype=np.float32) -> bytes:
    """PQTable::save (pq_table.rs:226-231)."""
    w = _W()
    w.u64(t.n_bits); w.u64(t.m); w.u32(DIST_TAG[t.dist]); w.opt_u64(t.k_means_size)
    w.u64(t.k_means_max_iter); w.f32(t.k_means_tol)
    w.u64(t.dim); w.u64(1 << t.n_bits)
    enc = (t.m + 1) // 2 if t.n_bits == 4 else t.m
    w.u64(enc)
    w.u64(enc); w.array(t.encoded_vec_set, np.uint8)          # VecSet<u8>{dim, data}
    w.u64(len(t.group_k_means))
    for km in t.group_k_means:
        _w_kmeans(w, km, dtype)
    w.array(t.dist_cache, np.float32)
    return bytes(w.b)


def load_pq_table(data, dtype=np.float32) -> PQTableRecord:
    r = _R(data)
    n_bits, m, dist, kms = r.u64(), r.u64(), DIST_NAME[r.u32()], r.opt_u64()
    it, tol, dim, k, enc = r.u64(), r.f32(), r.u64(), r.u64(), r.u64()
    if k != 1 << n_bits:
        raise ValueError("k != 2**n_bits")
    codes = _r_vecset(r, np.uint8)
    groups = [_r_kmeans(r, dtype) for _ in range(r.u64())]
    dc = r.array(np.float32)
    r.done()
    if codes.size and codes.shape[1] != enc:
        raise ValueError("encoded_dim mismatch")
    return PQTableRecord(n_bits, m, dist, kms, it, tol, dim, codes, groups, dc)


def pq_table_record(pq_table, codebooks=None) -> PQTableRecord:
    """Builds the on-disk record of a lab_1806_vec_db_b200.PQTable (device codes + host codebooks)."""
    from .index import pq_groups
    cfg = pq_table.config
    books = np.asarray(pq_table.codebooks if codebooks is None else codebooks)
    groups, off, kc = [], 0, pq_table.k
    dc = []
    for lo, hi in pq_groups(pq_table.dim, cfg.m):
        c = books[off:off + kc * (hi - lo)].reshape(kc, hi - lo)
        off += kc * (hi - lo)
        groups.append(KMeansRecord(kc, cfg.k_means_max_iter, cfg.k_means_tol, cfg.dist, (lo, hi), c))
        if cfg.dist == "cosine":   # sequential f32 dot(c, c), pq_table.rs:165-170
            dc.append(np.cumsum(c.astype(np.float32) * c.astype(np.float32), axis=1, dtype=np.float32)[:, -1])
        else:
            dc.append(np.zeros(kc, np.float32))
    return PQTableRecord(cfg.n_bits, cfg.m, cfg.dist, cfg.k_means_size, cfg.k_means_max_iter, cfg.k_means_tol,
                         pq_table.dim, pq_table.encoded_vec_set, groups, np.concatenate(dc))


class IVFIndexRecord(NamedTuple):
    dist: str
    default_n_probes: int
    vec_set: np.ndarray                   # [n, dim] (empty for save_without_vec_set)
    k: int
    k_means_size: Optional[int]
    k_means_max_iter: int
    k_means_tol: float
    clusters: List[np.ndarray]
    k_means: KMeansRecord


def dump_ivf_index(t: IVFIndexRecord, dtype=np.float32, dim=None) -> bytes:
    """IndexSerde::save for IVFIndex (mod.rs:122-127; ivf_index.rs:109-121 empties vec_set but keeps its dim)."""
    w = _W()
    w.u32(DIST_TAG[t.dist]); w.u64(t.default_n_probes)
    vs = np.asarray(t.vec_set, dtype)
    w.u64(vs.shape[1] if vs.ndim == 2 and vs.size else (dim or t.k_means.centroids.shape[1]))
    w.array(vs, dtype)
    w.u64(t.k); w.opt_u64(t.k_means_size); w.u64(t.k_means_max_iter); w.f32(t.k_means_tol)
    w.u64(len(t.clusters))
    for c in t.clusters:
        w.array(np.asarray(c, np.uint64), np.uint64)
    _w_kmeans(w, t.k_means, dtype)
    return bytes(w.b)


def load_ivf_index(data, dtype=np.float32) -> IVFIndexRecord:
    r = _R(data)
    dist, probes = DIST_NAME[r.u32()], r.u64()
    dim = r.u64()
    vdata = r.array(dtype)
    vs = vdata.reshape(-1, dim) if vdata.size else np.zeros((0, dim), dtype)
    k, kms, it, tol = r.u64(), r.opt_u64(), r.u64(), r.f32()
    clusters = [r.array(np.uint64) for _ in range(r.u64())]
    km = _r_kmeans(r, dtype)
    r.done()
    return IVFIndexRecord(dist, probes, vs, k, kms, it, tol, clusters, km)


def dump_flat_index(dist) -> bytes:
    """FlatIndex::save_without_vec_set writes only the distance tag (flat_index.rs:72-77)."""
    return struct.pack("<I", DIST_TAG[dist])


def load_flat_index(data) -> str:
    if len(data) != 4:
        raise ValueError("a Flat index-only file is exactly 4 bytes")
    return DIST_NAME[struct.unpack("<I", data)[0]]


# ---- HNSWIndex (hnsw_index.rs:99-141; serde field order = declaration order, dist_cache is #[serde(skip)]) ------------
class HNSWIndexRecord(NamedTuple):
    dim: int
    dist: str
    max_elements: int
    m: int
    max_m0: int
    ef_construction: int
    default_ef: int
    inv_log_m: float
    start_batch_since: int
    vec_set: np.ndarray          # [n, dim]; empty when saved with save_without_vec_set (:642-655)
    level0_links: np.ndarray     # [n * max_m0] u32
    other_links: list            # per node: [level * m] u32
    links_len: list              # per node: [level + 1] usize
    vec_level: np.ndarray        # [n] usize
    num_deleted: int
    enter_level: int             # or None
    enter_point: int             # or None


def dump_hnsw_index(t: HNSWIndexRecord, dtype=np.float32) -> bytes:
    w = _W()
    w.u64(t.dim); w.u32(DIST_TAG[t.dist]); w.u64(t.max_elements); w.u64(t.m); w.u64(t.max_m0)
    w.u64(t.ef_construction); w.u64(t.default_ef); w.f32(t.inv_log_m); w.u64(t.start_batch_since)
    vs = np.asarray(t.vec_set, dtype)
    w.u64(t.dim)
    w.array(vs, dtype)
    w.array(t.level0_links, np.uint32)
    w.u64(len(t.other_links))
    for v in t.other_links:
        w.array(v, np.uint32)
    w.u64(len(t.links_len))
    for v in t.links_len:
        w.array(v, np.uint64)
    w.array(t.vec_level, np.uint64)
    w.u64(t.num_deleted); w.opt_u64(t.enter_level); w.opt_u64(t.enter_point)
    return bytes(w.b)


def load_hnsw_index(data, dtype=np.float32) -> HNSWIndexRecord:
    r = _R(data)
    dim, dist, max_elements, m, max_m0 = r.u64(), DIST_NAME[r.u32()], r.u64(), r.u64(), r.u64()
    efc, default_ef, inv_log_m, sbs = r.u64(), r.u64(), r.f32(), r.u64()
    vdim = r.u64()
    vdata = r.array(dtype)
    vs = vdata.reshape(-1, vdim) if vdata.size else np.zeros((0, vdim), dtype)
    level0 = r.array(np.uint32)
    other = [r.array(np.uint32) for _ in range(r.u64())]
    lens = [r.array(np.uint64) for _ in range(r.u64())]
    vec_level = r.array(np.uint64)
    num_deleted, enter_level, enter_point = r.u64(), r.opt_u64(), r.opt_u64()
    r.done()
    return HNSWIndexRecord(dim, dist, max_elements, m, max_m0, efc, default_ef, inv_log_m, sbs, vs, level0, other, lens,
                           vec_level, num_deleted, enter_level, enter_point)


def hnsw_index_record(index, with_vec_set=False, rows=None) -> HNSWIndexRecord:
    """Record of a device-resident HNSWIndex (index.py) in the reference's layout."""
    links0, len0 = index.level0_links()
    levels, ulinks, ulen = index.upper_links()
    m = index.config.M
    other, lens, o = [], [], 0
    for i, lv in enumerate(levels):
        lv = int(lv)
        other.append(ulinks[o * m:(o + lv) * m].copy())
        lens.append(np.concatenate([[len0[i]], ulen[o:o + lv]]).astype(np.uint64))
        o += lv
    ep, el = index.enter_point
    n = len(levels)
    dist = "l2sqr" if index.vec_set.metric == 0 else "cosine"
    return HNSWIndexRecord(index.vec_set.dim, dist, index.config.max_elements, m, 2 * m, index.ef_construction,
                           index.default_ef, float(np.float32(1.0) / np.log(np.float32(m))), 1000,
                           (np.asarray(rows) if with_vec_set else np.zeros((0, index.vec_set.dim), index.vec_set.dtype)),
                           links0.reshape(-1), other, lens, np.asarray(levels, np.uint64), 0,
                           (el if n else None), (ep if n else None))


# ---- bench driver files (examples/bench.rs) --------------------------------------------------------------------------
def load_bench_config(path) -> dict:
    """BenchConfig (bench.rs:70-92) as a dict; `ef` is expanded like BenchEf (range -> list)."""
    import tomllib
    with open(path, "rb") as f:
        cfg = tomllib.load(f)
    for key in ("dist", "label", "gnd_path", "index_cache", "ef", "algorithm", "base", "test", "bench_output"):
        if key not in cfg:
            raise ValueError(f"missing field `{key}`")
    ef = cfg["ef"]
    if "range" in ef:
        r = ef["range"]
        cfg["ef_values"] = list(range(r["start"], r["end"] + 1, r["step"]))
    else:
        cfg["ef_values"] = list(ef.get("list", ef.get("values", [])))
    cfg["dist"] = DIST_TOML[cfg["dist"]]
    return cfg


def dump_result_list(title, results) -> str:
    """ResultList (bench.rs:312-368): results = [{label, ef: [...], search_time: [...ms], recall: [...]}, ...]."""
    out = [f'title = "{title}"', ""]
    for res in results:
        out.append("[[results]]")
        out.append(f'label = "{res["label"]}"')
        out.append("ef = [" + ", ".join(str(int(e)) for e in res["ef"]) + "]")
        out.append("search_time = [\n" + "".join(f"    {float(t)!r},\n" for t in res["search_time"]) + "]")
        out.append("recall = [\n" + "".join(f"    {float(t)!r},\n" for t in res["recall"]) + "]")
        out.append("")
    return "\n".join(out)


def load_result_list(text) -> dict:
    import tomllib
    d = tomllib.loads(text)
    d.setdefault("title", "")
    d.setdefault("results", [])
    return d
