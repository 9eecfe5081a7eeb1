"""Independent numpy-float32 emulation of the reference's sequential-f32 semantics.

Used ONLY to generate / re-check golden vectors (tests/golden/make_golden.py and the
`not gpu` tests). It deliberately shares no code with oracle/oracle.cpp so that the two
pin each other. Every reduction is `np.cumsum(..., dtype=float32)[..., -1]`, i.e. a strictly
left-to-right f32 sum of separately rounded f32 products, which is what rustc emits for
`a.iter().zip(b).map(|(x, y)| x * y).sum()` (reference src/distance/mod.rs:72-77).
"""
import bisect
import numpy as np

F = np.float32


def seq_sum(x):
    x = np.asarray(x, dtype=F)
    if x.shape[-1] == 0:
        return np.zeros(x.shape[:-1], dtype=F)
    return np.cumsum(x, axis=-1, dtype=F)[..., -1]


def dot(a, b):  # distance/mod.rs:72-74, :80-85
    return seq_sum(np.asarray(a, F) * np.asarray(b, F))


def l2sqr(a, b):  # distance/mod.rs:75-77, :86-94
    d = np.asarray(a, F) - np.asarray(b, F)
    return seq_sum(d * d)


def vec_norm(a):  # distance/mod.rs:46-48
    return np.sqrt(dot(a, a)).astype(F)


def cosine_cached(a, b, na, nb):  # distance/mod.rs:67-69
    den = np.maximum((np.asarray(na, F) * np.asarray(nb, F)).astype(F), F(1e-10))
    return (F(1.0) - (dot(a, b) / den).astype(F)).astype(F)


def cosine(a, b):  # distance/mod.rs:60-64
    return cosine_cached(a, b, vec_norm(a), vec_norm(b))


def l2sqr_cached(a, b, ipa, ipb):  # distance/mod.rs:54-57
    s = (np.asarray(ipa, F) + np.asarray(ipb, F)).astype(F)
    t = (F(2.0) * dot(a, b)).astype(F)
    return (s - t).astype(F)


def distance(a, b, metric):
    return l2sqr(a, b) if metric == "l2sqr" else cosine(a, b)


class ResultSet:
    """candidate_pair.rs:43-82 (BTreeSet of (distance, index), strict-< replacement)."""

    def __init__(self, k):
        self.k = k
        self.items = []  # sorted list of (distance, index)

    def add(self, d, i):
        d = float(d)
        key = (d, int(i))
        if len(self.items) < self.k:
            pos = bisect.bisect_left(self.items, key)
            if pos < len(self.items) and self.items[pos] == key:
                return True
            self.items.insert(pos, key)
            return True
        if self.items and d < self.items[-1][0]:
            self.items.pop()
            bisect.insort(self.items, key)
            return True
        return False


def flat_knn(base, queries, k, metric):
    """flat_index.rs:48-57: k smallest by (distance, index) (ascending scan + strict <)."""
    n = base.shape[0]
    kk = min(k, n)
    ids = np.zeros((len(queries), kk), np.int64)
    dd = np.zeros((len(queries), kk), F)
    for qi, q in enumerate(queries):
        d = distance(np.broadcast_to(q, base.shape), base, metric)
        order = np.lexsort((np.arange(n), d))[:kk]
        ids[qi] = order
        dd[qi] = d[order]
    return ids, dd


def pq_groups(dim, m):  # pq_table.rs:38-53
    out, cur = [], 0
    while cur < dim:
        rem = m - len(out)
        gs = -(-(dim - cur) // rem)
        out.append((cur, cur + gs))
        cur += gs
    return out


def find_nearest(v, centroids, metric):  # k_means.rs:40-57 (ties -> lowest id)
    d = distance(np.broadcast_to(v, centroids.shape), centroids, metric)
    return int(np.lexsort((np.arange(len(d)), d))[0])


def assign(rows, centroids, metric, lo=0, hi=None):  # k_means.rs:117-120
    hi = rows.shape[1] if hi is None else hi
    out = np.zeros(len(rows), np.uint32)
    for i, r in enumerate(rows):
        out[i] = find_nearest(r[lo:hi], centroids, metric)
    return out


def pq_encode(rows, codebooks, m, n_bits, metric):  # pq_table.rs:66-91
    """codebooks: list of [kc, len_g] arrays."""
    groups = pq_groups(rows.shape[1], m)
    idx = np.zeros((len(rows), m), np.int64)
    for g, (lo, hi) in enumerate(groups):
        idx[:, g] = assign(rows, codebooks[g], metric, lo, hi)
    if n_bits == 8:
        return idx.astype(np.uint8)
    enc = (m + 1) // 2
    codes = np.zeros((len(rows), enc), np.uint8)
    for i in range(m // 2):
        codes[:, i] = idx[:, 2 * i] | (idx[:, 2 * i + 1] << 4)
    if m % 2 == 1:
        codes[:, m // 2] = idx[:, m - 1]
    return codes


def pq_lookup(q, codebooks, m, metric):  # pq_table.rs:195-224
    groups = pq_groups(len(q), m)
    lut = []
    for g, (lo, hi) in enumerate(groups):
        c = codebooks[g]
        sub = np.broadcast_to(q[lo:hi], c.shape)
        lut.append(l2sqr(sub, c) if metric == "l2sqr" else dot(sub, c))
    qcache = F(0.0) if metric == "l2sqr" else vec_norm(q)
    return np.concatenate(lut).astype(F), F(qcache)


def pq_dist_cache(codebooks, metric):  # pq_table.rs:165-170
    return np.concatenate(
        [np.zeros(len(c), F) if metric == "l2sqr" else dot(c, c) for c in codebooks]
    ).astype(F)


def split_indices(codes, m, n_bits):  # pq_table.rs:55-65
    if n_bits == 8:
        return codes.astype(np.int64)
    lo = (codes & 0xF).astype(np.int64)
    hi = (codes >> 4).astype(np.int64)
    return np.stack([lo, hi], axis=-1).reshape(codes.shape[0], -1)[:, :m]


def pq_adc(codes, m, n_bits, lut, dist_cache, qcache, metric):  # pq_table.rs:239-301
    kc = 1 << n_bits
    idx = split_indices(codes, m, n_bits) + np.arange(m)[None, :] * kc
    s = seq_sum(lut[idx])
    if metric == "l2sqr":
        return s
    cdp = seq_sum(dist_cache[idx])
    den = np.maximum((np.sqrt(cdp).astype(F) * F(qcache)).astype(F), F(1e-10))
    return (F(1.0) - (s / den).astype(F)).astype(F)


def flat_knn_pq(base, codes, codebooks, m, n_bits, q, k, ef, metric):
    """flat_index.rs:84-104 + candidate_pair.rs:102-108."""
    lut, qcache = pq_lookup(q, codebooks, m, metric)
    dc = pq_dist_cache(codebooks, metric)
    adc = pq_adc(codes, m, n_bits, lut, dc, qcache, metric)
    kk = max(ef, k)
    order = np.lexsort((np.arange(len(adc)), adc))[:kk]  # ascending scan => k smallest by (d, i)
    rs = ResultSet(k)
    for i in order:
        rs.add(distance(q, base[i], metric), i)
    return rs.items, order, adc[order]


def find_n_nearest(q, centroids, n_probes, metric):  # k_means.rs:174-191
    d = distance(np.broadcast_to(q, centroids.shape), centroids, metric)
    return np.lexsort((np.arange(len(d)), d))[:n_probes]


def ivf_knn(base, centroids, lists, q, k, n_probes, metric):  # ivf_index.rs:143-154
    rs = ResultSet(k)
    for c in find_n_nearest(q, centroids, n_probes, metric):
        for i in lists[c]:
            rs.add(distance(base[i], q, metric), i)
    return rs.items
