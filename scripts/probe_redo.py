"""Which queries of the bench workload fail the tensor path's completeness check, and why (overflow / threshold)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
import lab_1806_vec_db_b200 as V
from lab_1806_vec_db_b200 import _lib as L
from lab_1806_vec_db_b200.sharded import unpack_keys
lib = L.lib(); dev = torch.device("cuda:0")
b1000, t1000 = B.load_fixtures()
n, nq, k = 1_000_000, 10_000, 100
base = B.synth(b1000, 0, n, 42, dev); q = B.synth(t1000, 0, nq, 43, dev)
vs = V.DeviceVecSet.from_device(base.data_ptr(), n, 960, 960, np.float32, "l2sqr", keepalive=base)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
nn, ns, mn, me = C.c_uint64(0), C.c_uint32(0), C.c_float(0), C.c_float(0)
L.check(lib.vdb_tq_info(vs._h, C.byref(nn), C.byref(ns), C.byref(mn), C.byref(me)))
j0 = int(lib.vdb_tq_j0(k, ns.value, n)); j = int(lib.vdb_tq_sample_j(j0, ns.value))
print("sample", ns.value, "j0", j0, "j", j)
tq = C.c_void_p(); L.check(lib.vdb_tq_begin_dev(vs._h, C.c_void_p(q.data_ptr()), nq, st, C.byref(tq)))
jk = torch.empty((nq, j), dtype=torch.int64, device=dev); L.check(lib.vdb_tq_sample_dev(tq, j, C.c_void_p(jk.data_ptr())))
tau = torch.empty((nq,), dtype=torch.float32, device=dev); L.check(lib.vdb_tq_tau_dev(tq, C.c_void_p(jk.data_ptr()), 1, j, min(j0, j), C.c_void_p(tau.data_ptr())))
keys = torch.empty((nq, k), dtype=torch.int64, device=dev); ovf = torch.empty((nq,), dtype=torch.int32, device=dev)
L.check(lib.vdb_tq_filter_dev(tq, k, C.c_void_p(tau.data_ptr()), C.c_void_p(keys.data_ptr()), C.c_void_p(ovf.data_ptr())))
redo = torch.empty((nq,), dtype=torch.int32, device=dev); nredo = torch.zeros((1,), dtype=torch.int32, device=dev)
L.check(lib.vdb_tq_check_dev(tq, C.c_void_p(keys.data_ptr()), k, n, C.c_void_p(tau.data_ptr()), C.c_void_p(ovf.data_ptr()), C.c_void_p(redo.data_ptr()), C.c_void_p(nredo.data_ptr())))
torch.cuda.synchronize()
nr = int(nredo.item()); sel = redo[:nr].cpu().numpy()
print("redo", nr, sel)
qsq = (q.double() ** 2).sum(1).cpu().numpy()
dk, _ = unpack_keys(keys.cpu().numpy().astype(np.uint64)); dj, _ = unpack_keys(jk.cpu().numpy().astype(np.uint64))
for qi in sel:
    print("q", qi, "ovf", int(ovf[qi]), "tau", float(tau[qi]), "d_k - qsq", float(dk[qi, k - 1]) - qsq[qi], "d_k", float(dk[qi, k - 1]),
          "sample exact d (first j):", dj[qi][:j], "valid keys", int((keys[qi] != -1).sum()))
lib.vdb_tq_end(tq)
