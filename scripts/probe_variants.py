"""Flat search on the other VecSet variants of the reference (u8 rows, cosine metric) at 1M x 960:
streaming scan (nq = 1, 8) and the batched tensor path (nq = 10000, k = 100). Device-resident, CUDA events."""
import ctypes as C
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lab_1806_vec_db_b200 as V
from lab_1806_vec_db_b200 import _lib as L
from bench import synth, load_fixtures, DIM

n = 1_000_000
dev = torch.device("cuda:0")
b1000, t1000 = load_fixtures()
base = synth(b1000, 0, n, 42, dev)
qf = synth(t1000, 0, 10_000, 43, dev)
lib = L.lib()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for name, rows, q, dtype, metric in (("f32 cosine", base, qf, np.float32, "cosine"),
                                     ("u8 l2sqr", (base * 255).round().clamp(0, 255).to(torch.uint8), (qf * 255).round().clamp(0, 255).to(torch.uint8), np.uint8, "l2sqr"),
                                     ("u8 cosine", (base * 255).round().clamp(0, 255).to(torch.uint8), (qf * 255).round().clamp(0, 255).to(torch.uint8), np.uint8, "cosine")):
    ds = V.DeviceVecSet.from_device(rows.data_ptr(), n, DIM, DIM, dtype, metric, keepalive=rows)
    esz = 4 if dtype == np.float32 else 1
    for nq, k in ((1, 10), (8, 10), (10_000, 100)):
        qq = q[:nq].contiguous()
        ids = torch.empty((nq, k), dtype=torch.int64, device=dev); dd = torch.empty((nq, k), dtype=torch.float32, device=dev)
        cnt = torch.empty(nq, dtype=torch.int32, device=dev)
        def run():
            L.check(lib.vdb_flat_knn_dev(ds._h, C.c_void_p(qq.data_ptr()), nq, k, C.c_void_p(ids.data_ptr()),
                                         C.c_void_p(dd.data_ptr()), C.c_void_p(cnt.data_ptr()), st))
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        extra = f"{n * DIM * esz / ms / 1e6:7.0f} GB/s of rows" if nq <= 8 else f"{2.0 * nq * n * DIM / ms / 1e9:6.0f} TFLOP/s"
        print(f"{name:10s} nq={nq:5d} k={k:3d}: {ms:8.3f} ms  {nq / ms * 1e3:10.0f} QPS  {extra}", flush=True)
    g0, g1, g2 = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    L.check(lib.vdb_flat_gemm_stats(C.byref(g0), C.byref(g1), C.byref(g2)))
    print(f"{name}: tensor path cumulative queries {g0.value}, candidates/query {g1.value / max(g0.value, 1):.0f}, exact fallbacks {g2.value}", flush=True)
    ds.close()
