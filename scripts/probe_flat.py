"""Quick device-side timing of the Flat scan (K1) at several batch sizes. Not the bench."""
import ctypes as C
import sys, os
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lab_1806_vec_db_b200 as V
from lab_1806_vec_db_b200 import _lib as L

n = int(os.environ.get("N", 1_000_000)); dim = int(os.environ.get("DIM", 960))
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(42)
u8 = os.environ.get("DTYPE", "f32") == "u8"
base = torch.rand((n, dim), device=dev, generator=g, dtype=torch.float32)
if u8:
    base = (base * 255).round().to(torch.uint8)
esz = 1 if u8 else 4
ds = V.DeviceVecSet.from_device(base.data_ptr(), n, dim, dim, np.uint8 if u8 else np.float32, os.environ.get("METRIC", "l2sqr"), keepalive=base)
lib = L.lib()
L.check(lib.vdb_flat_set_path(int(os.environ.get("FLATPATH", 1))))
stream = torch.cuda.current_stream().cuda_stream
cases = [(1, 10), (2, 10), (4, 10), (8, 10), (8, 100), (1, 100), (16, 10), (64, 100)]
if os.environ.get("CASES"):  # e.g. CASES=8:10,4:10
    cases = [tuple(int(v) for v in c.split(":")) for c in os.environ["CASES"].split(",")]
for nq, k in cases:
    q = torch.rand((nq, dim), device=dev, generator=g)
    if u8:
        q = (q * 255).round().to(torch.uint8)
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev); dd = torch.empty((nq, k), dtype=torch.float32, device=dev)
    cnt = torch.empty(nq, dtype=torch.int32, device=dev)
    def run():
        L.check(lib.vdb_flat_knn_dev(ds._h, C.c_void_p(q.data_ptr()), nq, k, C.c_void_p(ids.data_ptr()),
                                     C.c_void_p(dd.data_ptr()), C.c_void_p(cnt.data_ptr()), C.c_void_p(stream)))
    for _ in range(int(os.environ.get("WARM", 3))): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = int(os.environ.get("REPS", 10))
    e0.record()
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    passes = (nq + 7) // 8
    gbs = passes * n * dim * esz / ms / 1e6
    extra = ""
    if os.environ.get("PROF"):
        L.check(lib.vdb_prof_reset()); L.check(lib.vdb_prof_enable(1))
        for _ in range(50): run()
        torch.cuda.synchronize()
        L.check(lib.vdb_prof_enable(0))
        t, c = C.c_double(0), C.c_uint64(0)
        L.check(lib.vdb_prof_read(b"flat_scan", C.byref(t), C.byref(c)))
        extra = f"  scan kernel {t.value / max(1, c.value) * 1e3:.1f} us x {c.value // 50} per call"
    print(f"nq={nq:3d} k={k:4d}: {ms:8.3f} ms/call  {nq/ms*1e3:9.1f} QPS  {gbs:7.1f} GB/s ({gbs/6551.4*100:5.1f}% of measured HBM peak){extra}", flush=True)
