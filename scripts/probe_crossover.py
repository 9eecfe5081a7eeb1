"""Scan vs tensor path per batch size (device-resident), to place the auto-path crossover."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lab_1806_vec_db_b200 as V
from lab_1806_vec_db_b200 import _lib as L
from bench import synth, load_fixtures, DIM
n = 1_000_000
dev = torch.device("cuda:0")
b1000, t1000 = load_fixtures()
base = synth(b1000, 0, n, 42, dev)
q_all = synth(t1000, 0, 1024, 43, dev)
ds = V.DeviceVecSet.from_device(base.data_ptr(), n, DIM, DIM, np.float32, "l2sqr", keepalive=base)
lib = L.lib()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for k in (10, 100):
    for nq in (4, 5, 6, 8, 12, 16, 32, 64, 128, 129, 256, 512, 1024):
        q = q_all[:nq].contiguous()
        ids = torch.empty((nq, k), dtype=torch.int64, device=dev); dd = torch.empty((nq, k), dtype=torch.float32, device=dev)
        cnt = torch.empty(nq, dtype=torch.int32, device=dev)
        res = {}
        for path in (1, 2):
            L.check(lib.vdb_flat_set_path(path))
            def run():
                L.check(lib.vdb_flat_knn_dev(ds._h, C.c_void_p(q.data_ptr()), nq, k, C.c_void_p(ids.data_ptr()),
                                             C.c_void_p(dd.data_ptr()), C.c_void_p(cnt.data_ptr()), st))
            for _ in range(2): run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record()
            for _ in range(reps): run()
            e1.record(); torch.cuda.synchronize()
            res[path] = e0.elapsed_time(e1) / reps
        print(f"k={k:3d} nq={nq:5d}: scan {res[1]:8.3f} ms ({nq/res[1]*1e3:9.0f} QPS)   tensor {res[2]:8.3f} ms ({nq/res[2]*1e3:9.0f} QPS)", flush=True)
q, c, f = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
lib.vdb_flat_gemm_stats(C.byref(q), C.byref(c), C.byref(f)); print("tensor stats: queries", q.value, "cands/query", c.value / max(q.value, 1), "fallbacks", f.value)
