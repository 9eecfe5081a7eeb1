"""Calibration of the second synthetic set (bench.synth_clustered): recall@10 of IVF / PQ / HNSW on 1M x 960 for a few
(sub-clusters, spread, noise) settings, looking for the reference's published operating range 0.85-0.95."""
import ctypes as C, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
import lab_1806_vec_db_b200 as V
from lab_1806_vec_db_b200 import _lib as L
from lab_1806_vec_db_b200.index import train_codebooks
lib = L.lib(); dev = torch.device("cuda:0")
b1000, t1000 = B.load_fixtures()
n, nq, k = int(os.environ.get("N", 1_000_000)), 1000, 10
def rec(ids, gt): return float(np.mean([len(set(a) & set(b)) / k for a, b in zip(ids.tolist(), gt.tolist())]))
for (latent, spread, noise) in [(32, 0.012, 0.017), (32, 0.010, 0.018), (32, 0.008, 0.019), (32, 0.006, 0.02)]:
    base = B.synth_clustered(b1000, 0, n, 42, dev, latent, spread, noise)
    # queries: members of the same mixture (fresh noise), like GIST's query set is drawn from the base distribution
    q = B.synth_clustered(b1000, 0, nq, 43, dev, latent, spread, noise)
    vs = V.DeviceVecSet.from_device(base.data_ptr(), n, 960, 960, np.float32, "l2sqr", keepalive=base)
    flat = V.FlatIndex(vs)
    qh = q.cpu().numpy()
    gt = flat.knn_batch(qh, k)
    row = {"latent": latent, "spread": spread, "noise": noise, "d1": float(gt[1][:, 0].mean()), "d10": float(gt[1][:, 9].mean())}
    rng = np.random.default_rng(42)
    bh = None
    sel = torch.as_tensor(rng.permutation(n)[:100_000], device=dev)
    km = V.KMeans.from_vec_set(np.ascontiguousarray(base.index_select(0, sel).cpu().numpy()), V.KMeansConfig(128, 20, 1e-6, "l2sqr"), rng)
    ivf = V.IVFIndex(vs, km.centroids)
    for nprobe in (8, 24):
        row[f"ivf{nprobe}"] = rec(ivf.knn_with_ef_batch(qh, k, nprobe)[0], gt[0])
    del ivf
    cfg = V.PQConfig(4, 240, "l2sqr", 10_000, 20, 1e-6)
    sel = torch.as_tensor(rng.permutation(n)[:10_000], device=dev)
    td = V.DeviceVecSet(np.ascontiguousarray(base.index_select(0, sel).cpu().numpy()), "l2sqr")
    books = train_codebooks(td, cfg, rng); td.close()
    pq = V.PQTable(vs, cfg, books)
    for ef in (240, 600):
        row[f"pq{ef}"] = rec(flat.knn_pq_batch(qh, k, ef, pq)[0], gt[0])
    t0 = time.perf_counter()
    hn = V.HNSWIndex(vs, V.HNSWConfig(0, 200, 16), rng=np.random.default_rng(42))
    row["hnsw_build_s"] = time.perf_counter() - t0
    for ef in (120, 360):
        row[f"hnsw{ef}"] = rec(hn.knn_with_ef_batch(qh, k, ef)[0], gt[0])
    for ef in (240, 600):
        row[f"hnswpq{ef}"] = rec(hn.knn_pq_batch(qh, k, ef, pq)[0], gt[0])
    print(json.dumps(row), flush=True)
    del hn, pq, flat, vs, base
    torch.cuda.empty_cache()
