"""Row-part sweep of the tensor filter pass (VDB_GEMM_PARTS x VDB_GEMM_PART_RATIO) on the bench workload, plus the
small-batch tensor pass (nq = 1 .. 128) against the exact scan. Device-resident, CUDA events."""
import ctypes as C, os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
import lab_1806_vec_db_b200 as V
from lab_1806_vec_db_b200 import _lib as L
lib = L.lib()
dev = torch.device("cuda:0")
b1000, t1000 = B.load_fixtures()
n = int(os.environ.get("N", 1_000_000))
base = B.synth(b1000, 0, n, 42, dev); q = B.synth(t1000, 0, 10_000, 43, dev)
vs = V.DeviceVecSet.from_device(base.data_ptr(), n, 960, 960, np.float32, "l2sqr", keepalive=base)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(qq, k, reps):
    nq = qq.shape[0]
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev); dd = torch.empty((nq, k), dtype=torch.float32, device=dev); cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
    f = lambda: L.check(lib.vdb_flat_knn_dev(vs._h, C.c_void_p(qq.data_ptr()), nq, k, C.c_void_p(ids.data_ptr()), C.c_void_p(dd.data_ptr()), C.c_void_p(cnt.data_ptr()), st))
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, ids
out = {"parts": [], "small": []}
for parts, ratio in [(1, 1.0), (3, 1.0), (3, 0.6), (4, 0.6), (4, 0.5), (5, 0.6), (6, 0.6), (6, 0.7), (8, 0.7)]:
    os.environ["VDB_GEMM_PARTS"] = str(parts); os.environ["VDB_GEMM_PART_RATIO"] = str(ratio)
    ms, _ = run(q, 100, 5)
    out["parts"].append((parts, ratio, round(ms, 3)))
    print("parts", parts, ratio, round(ms, 3), flush=True)
del os.environ["VDB_GEMM_PARTS"]; del os.environ["VDB_GEMM_PART_RATIO"]
for nq in (1, 2, 3, 4, 5, 8, 16, 32, 64, 128, 256, 1000):
    qq = q[:nq].contiguous()
    row = {"nq": nq}
    for name, path in (("scan", 1), ("tensor", 2)):
        if name == "scan" and nq > 16: continue
        L.check(lib.vdb_flat_set_path(path))
        ms, ids = run(qq, 10, 20)
        row[name] = round(ms, 4)
        row[name + "_ids"] = ids.cpu().numpy()
    if "scan" in row: row["same"] = bool((row["scan_ids"] == row["tensor_ids"]).all())
    row.pop("scan_ids", None); row.pop("tensor_ids", None)
    out["small"].append(row); print(row, flush=True)
L.check(lib.vdb_flat_set_path(0))
print(json.dumps(out))
