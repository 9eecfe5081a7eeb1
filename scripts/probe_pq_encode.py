import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import lab_1806_vec_db_b200 as V
from lab_1806_vec_db_b200 import _lib as L
from bench import synth, load_fixtures, DIM
n = 1_000_000
dev = torch.device("cuda:0")
b1000, t1000 = load_fixtures()
base = synth(b1000, 0, n, 42, dev)
ds = V.DeviceVecSet.from_device(base.data_ptr(), n, DIM, DIM, np.float32, "l2sqr", keepalive=base)
host = base[:10000].cpu().numpy()
books = np.concatenate([np.ascontiguousarray(host[100:116, lo:hi]).reshape(-1) for lo, hi in V.pq_groups(DIM, 240)])
lib = L.lib()
for rep in range(2):
    L.check(lib.vdb_prof_reset()); L.check(lib.vdb_prof_enable(1))
    t0 = time.perf_counter()
    pq = V.PQTable(ds, V.PQConfig(4, 240, "l2sqr"), books)
    wall = time.perf_counter() - t0
    L.check(lib.vdb_prof_enable(0))
    t, c = C.c_double(0), C.c_uint64(0)
    L.check(lib.vdb_prof_read(b"pq_encode", C.byref(t), C.byref(c)))
    print(f"pq create wall {wall*1e3:.1f} ms, encode kernel {t.value:.2f} ms ({c.value} launches)")
    pq.close()
