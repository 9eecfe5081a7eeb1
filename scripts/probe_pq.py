"""Where does a batched Flat+PQ search (C4 shape) spend its time?"""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lab_1806_vec_db_b200 as V
from lab_1806_vec_db_b200 import _lib as L
from bench import synth, load_fixtures, DIM
n = int(os.environ.get("N", 1_000_000))
M = int(os.environ.get("M", 240))
dev = torch.device("cuda:0")
b1000, t1000 = load_fixtures()
base = synth(b1000, 0, n, 42, dev)
ds = V.DeviceVecSet.from_device(base.data_ptr(), n, DIM, DIM, np.float32, "l2sqr", keepalive=base)
rng = np.random.default_rng(42)
host = base.cpu().numpy()
train = np.ascontiguousarray(host[rng.permutation(n)[:10_000]])
td = V.DeviceVecSet(train, "l2sqr")
books = np.concatenate([V.KMeans.from_vec_set(td, V.KMeansConfig(16, 20, 1e-6, "l2sqr", (lo, hi)), rng).centroids.reshape(-1)
                        for lo, hi in V.pq_groups(DIM, M)])
pq = V.PQTable(ds, V.PQConfig(4, M, "l2sqr", 10_000, 20, 1e-6), books)
lib = L.lib()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
k = 10
names = (b"pq_adc", b"pq_gemm", b"pq_exact", b"merge", b"rerank", b"pq_lut")
for nq in (1000, 10000):
    qq = synth(t1000, 0, nq, 43, dev)
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev); dd = torch.empty((nq, k), dtype=torch.float32, device=dev)
    cnt = torch.empty(nq, dtype=torch.int32, device=dev)
    for ef in (240, 600):
        def run():
            L.check(lib.vdb_pq_knn_dev(ds._h, pq._h, C.c_void_p(qq.data_ptr()), nq, k, ef, C.c_void_p(ids.data_ptr()),
                                       C.c_void_p(dd.data_ptr()), C.c_void_p(cnt.data_ptr()), st))
        run(); torch.cuda.synchronize()
        L.check(lib.vdb_prof_reset()); L.check(lib.vdb_prof_enable(1))
        t0 = time.perf_counter()
        for _ in range(3): run()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / 3 * 1e3
        L.check(lib.vdb_prof_enable(0))
        out = {}
        for name in names:
            t, c = C.c_double(0), C.c_uint64(0)
            L.check(lib.vdb_prof_read(name, C.byref(t), C.byref(c)))
            out[name.decode()] = (round(t.value / 3, 2), c.value // 3)
        print(f"nq={nq} ef={ef}: wall {wall:.2f} ms ({nq/wall*1e3:.0f} QPS) kernels (ms/call, launches): {out}", flush=True)
