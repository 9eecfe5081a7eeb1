"""Row-sharded search on 2 GPUs (NCCL): global-threshold tensor path and scan path vs the unsharded exact scan.
Skipped when fewer than 2 GPUs are visible (the round-end `-m gpu` run has one)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, {root!r})
import lab_1806_vec_db_b200 as V
from lab_1806_vec_db_b200 import _lib as L
from lab_1806_vec_db_b200.sharded import ShardedFlatIndex, shard_bounds
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
L.check(L.lib().vdb_set_device(local))
rng = np.random.default_rng(0)
n, nq, dim = 300_000, 300, 960
proto = rng.random((256, dim), dtype=np.float32) * 0.15
base = (proto[rng.integers(0, 256, n)] + 0.02 * rng.standard_normal((n, dim), dtype=np.float32)).clip(0, 1).astype(np.float32)
q = (proto[rng.integers(0, 256, nq)] + 0.02 * rng.standard_normal((nq, dim), dtype=np.float32)).clip(0, 1).astype(np.float32)
base[7777] = base[123]                     # an exact duplicate across the shard boundary region
lo, hi = shard_bounds(n, world, rank)
vs = V.DeviceVecSet(np.ascontiguousarray(base[lo:hi]), "l2sqr", id_base=lo)
idx = ShardedFlatIndex(vs, rank, world)
qd = torch.from_numpy(q).to(dev)
L.check(L.lib().vdb_flat_set_path(1))
full = V.FlatIndex.from_vec_set(base, "l2sqr")   # unsharded exact scan on this rank's GPU
for k in (10, 100):
    want = full.knn_batch(q, k)
    ids, dd, cnt = idx.knn_batch_dev(qd, k)                     # tensor phases with global thresholds
    assert (ids.cpu().numpy() == want[0].astype(np.int64)).all(), ("tensor", k, rank)
    assert (dd.cpu().numpy().view(np.uint32) == want[1].view(np.uint32)).all()
    assert (cnt.cpu().numpy() == want[2]).all()
    ids, dd, cnt = idx.knn_batch_dev(qd[:5].contiguous(), k)   # small batch: sharded exact scan
    assert (ids.cpu().numpy() == want[0][:5].astype(np.int64)).all(), ("scan", k, rank)
    assert (dd.cpu().numpy().view(np.uint32) == want[1][:5].view(np.uint32)).all()
# ---- sharded IVF and PQ: identical to the unsharded calls ----
from lab_1806_vec_db_b200.sharded import ShardedIVFIndex, ShardedPQFlatIndex
L.check(L.lib().vdb_flat_set_path(0))
cent = np.ascontiguousarray(base[np.random.default_rng(1).permutation(n)[:32]])
ivf_full = V.IVFIndex(full.vec_set, cent)
sivf = ShardedIVFIndex(V.IVFIndex(vs, cent), rank, world)
for nprobe in (1, 4, 32):
    want = ivf_full.knn_with_ef_batch(q[:64], 10, nprobe)
    ids, dd, cnt = sivf.knn_with_ef_batch_dev(qd[:64].contiguous(), 10, nprobe)
    assert (ids.cpu().numpy() == want[0].astype(np.int64)).all(), ("ivf", nprobe, rank)
    assert (dd.cpu().numpy().view(np.uint32) == want[1].view(np.uint32)).all()
books = np.concatenate([np.ascontiguousarray(base[100:116, lo:hi]).reshape(-1) for lo, hi in V.pq_groups(dim, 240)])
cfg = V.PQConfig(4, 240, "l2sqr")
pq_full = V.PQTable(full.vec_set, cfg, books)
spq = ShardedPQFlatIndex(vs, V.PQTable(vs, cfg, books), rank, world)
for k, ef in ((10, 240), (10, 5), (3, 600)):
    want = full.knn_pq_batch(q[:32], k, ef, pq_full)
    ids, dd, cnt = spq.knn_pq_batch_dev(qd[:32].contiguous(), k, ef)
    assert (ids.cpu().numpy() == want[0].astype(np.int64)).all(), ("pq", k, ef, rank)
    assert (dd.cpu().numpy().view(np.uint32) == want[1].view(np.uint32)).all()
# ---- HNSW: replicas only, the query batch is split across the ranks ----
from lab_1806_vec_db_b200.sharded import ReplicatedHNSWIndex
sub = np.ascontiguousarray(base[:20000])
hn = V.HNSWIndex(V.DeviceVecSet(sub, "l2sqr"), V.HNSWConfig(0, 100, 12), levels=V.hnsw_rand_levels(20000, 12, np.random.default_rng(2)))
want = hn.knn_with_ef_batch(q[:101], 10, 64)
ids, dd, cnt = ReplicatedHNSWIndex(hn, rank, world).knn_with_ef_batch_dev(qd[:101].contiguous(), 10, 64)
exact = V.FlatIndex.from_vec_set(sub, "l2sqr").knn_batch(q[:101], 10)[0].astype(np.int64)
got = ids.cpu().numpy()
assert got.shape == (101, 10) and (cnt.cpu().numpy() == 10).all()
lo_q, hi_q = rank * 51, min(101, rank * 51 + 51)     # this rank's slice was searched on this rank's replica
assert (got[lo_q:hi_q] == want[0][lo_q:hi_q].astype(np.int64)).all()
assert np.mean([len(set(a) & set(b)) / 10 for a, b in zip(got.tolist(), exact.tolist())]) >= 0.9   # every rank's part is a real search
pin = torch.from_numpy(q).pin_memory()
out = idx.knn_batch(pin, 10)
assert (out[0].numpy() == full.knn_batch(q, 10)[0].astype(np.int64)).all()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_two_gpu_sharded_search(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-6000:]
    assert r.stdout.count("ok") == 2
