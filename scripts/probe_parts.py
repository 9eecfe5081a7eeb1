"""Filter pass cut into row parts (rerank of part i under the contraction of part i + 1): step time vs VDB_GEMM_PARTS.
Device-resident, CUDA events, bench workload (1M x 960, 10 000 queries, k = 100) or N / NQ from the environment."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lab_1806_vec_db_b200 as V
from lab_1806_vec_db_b200 import _lib as L
from bench import synth, load_fixtures, DIM
n = int(os.environ.get("N", 1_000_000)); nq = int(os.environ.get("NQ", 10_000)); k = int(os.environ.get("K", 100))
dev = torch.device("cuda:0")
b1000, t1000 = load_fixtures()
base = synth(b1000, 0, n, 42, dev)
q = synth(t1000, 0, nq, 43, dev)
ds = V.DeviceVecSet.from_device(base.data_ptr(), n, DIM, DIM, np.float32, "l2sqr", keepalive=base)
lib = L.lib()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
ids = torch.empty((nq, k), dtype=torch.int64, device=dev); dd = torch.empty((nq, k), dtype=torch.float32, device=dev)
cnt = torch.empty(nq, dtype=torch.int32, device=dev)
ref = None
for parts in (1, 2, 3, 4, 6, 8, 0):
    os.environ["VDB_GEMM_PARTS"] = str(parts)
    def run():
        L.check(lib.vdb_flat_knn_dev(ds._h, C.c_void_p(q.data_ptr()), nq, k, C.c_void_p(ids.data_ptr()),
                                     C.c_void_p(dd.data_ptr()), C.c_void_p(cnt.data_ptr()), st))
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 8
    e0.record()
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    same = ""
    if ref is None: ref = (ids.clone(), dd.clone())
    else: same = f"  identical to parts=1: {bool((ids == ref[0]).all() and (dd.view(torch.int32) == ref[1].view(torch.int32)).all())}"
    print(f"parts={parts} (0 = default): {ms:8.3f} ms/step  {nq / ms * 1e3:9.0f} QPS{same}", flush=True)
