"""Where does a batched IVF search spend its time (kernel vs host grouping)?"""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lab_1806_vec_db_b200 as V
from lab_1806_vec_db_b200 import _lib as L
from bench import synth, load_fixtures, DIM
n = 1_000_000
dev = torch.device("cuda:0")
b1000, t1000 = load_fixtures()
base = synth(b1000, 0, n, 42, dev)
q = synth(t1000, 0, 1000, 43, dev)
ds = V.DeviceVecSet.from_device(base.data_ptr(), n, DIM, DIM, np.float32, "l2sqr", keepalive=base)
rng = np.random.default_rng(42)
host = base.cpu().numpy()
km = V.KMeans.from_vec_set(np.ascontiguousarray(host[rng.permutation(n)[:100_000]]), V.KMeansConfig(128, 20, 1e-6, "l2sqr"), rng)
ivf = V.IVFIndex(ds, km.centroids)
lib = L.lib()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
k = 10
for nq in (1000, 10000):
    qq = q if nq == 1000 else synth(t1000, 0, nq, 43, dev)
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev); dd = torch.empty((nq, k), dtype=torch.float32, device=dev)
    cnt = torch.empty(nq, dtype=torch.int32, device=dev)
    for nprobe in (8, 24):
        def run():
            L.check(lib.vdb_ivf_knn_dev(ds._h, ivf._h, C.c_void_p(qq.data_ptr()), nq, k, nprobe, C.c_void_p(ids.data_ptr()),
                                        C.c_void_p(dd.data_ptr()), C.c_void_p(cnt.data_ptr()), st))
        run(); torch.cuda.synchronize()
        L.check(lib.vdb_prof_reset()); L.check(lib.vdb_prof_enable(1))
        t0 = time.perf_counter()
        for _ in range(3): run()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / 3 * 1e3
        L.check(lib.vdb_prof_enable(0))
        out = {}
        for name in (b"ivf_scan", b"kmeans_assign", b"merge", b"flat_gemm", b"rerank"):
            t, c = C.c_double(0), C.c_uint64(0)
            L.check(lib.vdb_prof_read(name, C.byref(t), C.byref(c)))
            out[name.decode()] = round(t.value / 3, 2)
        g0, g1, g2 = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        L.check(lib.vdb_flat_gemm_stats(C.byref(g0), C.byref(g1), C.byref(g2)))
        out["cum stats(queries,cands,fallbacks)"] = [g0.value, g1.value, g2.value]
        print(f"nq={nq} nprobe={nprobe}: wall {wall:.2f} ms ({nq/wall*1e3:.0f} QPS) kernels(ms/call): {out}", flush=True)
