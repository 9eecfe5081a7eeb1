"""Generate the committed golden fixtures under tests/golden/.

Run HERE (the container that has /root/reference): `python tests/golden/make_golden.py`.
The GPU box has no /root/reference, so tests only ever read the .npz files this writes.

Inputs  : /root/reference/data/gist_1000.bin, gist_test.bin (1000x960 f32 LE, headerless;
          reference config/gist_1000.toml:1-3).
Outputs : fixtures.npz  - the two data files, losslessly re-encoded as uint16 (every value
                          is a multiple of 1e-4; float32(k / 1e4) reproduces the exact
                          bits, which this script asserts) + their sha256.
          golden_*.npz  - expected results computed by np_emul.py (independent numpy
                          float32 emulation of the reference's sequential arithmetic).
The SHA-256 anchors printed at the end are the ones quoted in SURVEY.md section 8(c).
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import np_emul as E  # noqa: E402

REF = "/root/reference/data"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    base = np.fromfile(f"{REF}/gist_1000.bin", dtype=np.float32).reshape(1000, 960)
    test = np.fromfile(f"{REF}/gist_test.bin", dtype=np.float32).reshape(1000, 960)
    fx = {}
    for name, a in (("base", base), ("test", test)):
        k = np.rint(a.astype(np.float64) * 1e4).astype(np.uint16)
        back = (k.astype(np.float64) / 1e4).astype(np.float32)
        assert (back.view(np.uint32) == a.view(np.uint32)).all(), "u16 re-encoding is not lossless"
        fx[name + "_u16"] = k
        fx[name + "_sha256"] = np.array(sha(a))
    np.savez_compressed(f"{HERE}/fixtures.npz", **fx)

    out = {}
    # --- C1: Flat kNN k=10, all 1000 queries (gen_gnd.rs:54-68 protocol) -------------------
    for metric in ("l2sqr", "cosine"):
        ids, dd = E.flat_knn(base, test, 10, metric)
        out[f"flat_{metric}_ids"] = ids
        out[f"flat_{metric}_dist"] = dd
        print(metric, "ids sha256", sha(ids.astype("<i8")), "dist sha256", sha(dd.astype("<f4")))
    # --- reference unit-test shape: dim clipped to 12, query=row 200, k=6 ------------------
    b12 = np.ascontiguousarray(base[:, :12])
    for metric in ("l2sqr", "cosine"):
        ids, dd = E.flat_knn(b12, b12[200:201], 6, metric)
        out[f"unit12_{metric}_ids"] = ids
        out[f"unit12_{metric}_dist"] = dd
        print("unit12", metric, ids[0].tolist(), dd[0].tolist())
    # --- k-means assignment given centroids (k_means.rs:117-120; ivf_index.rs:89-93) -------
    cent = np.ascontiguousarray(base[7:7 + 16])
    for metric in ("l2sqr", "cosine"):
        out[f"assign16_{metric}"] = E.assign(base, cent, metric)
    out["assign16_sel_l2sqr"] = E.assign(base, np.ascontiguousarray(cent[:, 100:113]), "l2sqr", 100, 113)
    # --- IVF probe scan given centroids/lists (ivf_index.rs:89-96, 143-154) ------------------
    for metric in ("l2sqr", "cosine"):
        a = out[f"assign16_{metric}"]
        lists = [np.nonzero(a == c)[0] for c in range(16)]
        ids = np.zeros((50, 10), np.int64)
        dd = np.zeros((50, 10), np.float32)
        for qi in range(50):
            items = E.ivf_knn(base, cent, lists, test[qi], 10, 4, metric)
            ids[qi] = [i for _, i in items]
            dd[qi] = [d for d, _ in items]
        out[f"ivf_{metric}_ids"], out[f"ivf_{metric}_dist"] = ids, dd
    # --- PQ: m=240 4-bit (config/bench_pq_240_hnsw.toml:16-23 shape) and m=7 8-bit/odd 4-bit --
    for tag, m, n_bits, dimclip in (("pq240", 240, 4, 960), ("pq7", 7, 4, 13), ("pq5b8", 5, 8, 13)):
        rows = np.ascontiguousarray(base[:200, :dimclip])
        kc = 1 << n_bits
        groups = E.pq_groups(dimclip, m)
        # deterministic codebooks: centroid c of group g = sub-vector of base row (3*c + g) % 1000
        cbs = [np.ascontiguousarray(
            base[[(3 * c + g) % 1000 for c in range(kc)], lo:hi]) for g, (lo, hi) in enumerate(groups)]
        out[f"{tag}_codebooks"] = np.concatenate([c.reshape(-1) for c in cbs])
        for metric in ("l2sqr", "cosine"):
            codes = E.pq_encode(rows, cbs, m, n_bits, metric)
            out[f"{tag}_{metric}_codes"] = codes
            dc = E.pq_dist_cache(cbs, metric)
            luts, adcs, qcs = [], [], []
            knn_ids = np.zeros((5, 10), np.int64)
            knn_dd = np.zeros((5, 10), np.float32)
            for qi in range(5):
                q = np.ascontiguousarray(test[qi, :dimclip])
                lut, qc = E.pq_lookup(q, cbs, m, metric)
                luts.append(lut)
                qcs.append(qc)
                adcs.append(E.pq_adc(codes, m, n_bits, lut, dc, qc, metric))
                items, _, _ = E.flat_knn_pq(rows, codes, cbs, m, n_bits, q, 10, 40, metric)
                knn_ids[qi] = [i for _, i in items]
                knn_dd[qi] = [d for d, _ in items]
            out[f"{tag}_{metric}_lut"] = np.stack(luts)
            out[f"{tag}_{metric}_qcache"] = np.array(qcs, np.float32)
            out[f"{tag}_{metric}_adc"] = np.stack(adcs)
            out[f"{tag}_{metric}_dist_cache"] = dc
            out[f"{tag}_{metric}_knn_ids"] = knn_ids
            out[f"{tag}_{metric}_knn_dist"] = knn_dd
    # --- cached-form distances (hnsw_index.rs:351-355; distance/mod.rs:54-57, 67-69) ----------
    cand = np.arange(0, 1000, 7)
    for metric in ("l2sqr", "cosine"):
        cache = E.dot(base, base) if metric == "l2sqr" else E.vec_norm(base)
        q = test[3]
        qc = E.dot(q, q) if metric == "l2sqr" else E.vec_norm(q)
        qq = np.broadcast_to(q, base[cand].shape)
        d = (E.l2sqr_cached(qq, base[cand], qc, cache[cand]) if metric == "l2sqr"
             else E.cosine_cached(qq, base[cand], qc, cache[cand]))
        out[f"cached_{metric}_dist"] = d
        out[f"cached_{metric}_rowcache"] = cache.astype(np.float32)
    np.savez_compressed(f"{HERE}/golden.npz", **out)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
