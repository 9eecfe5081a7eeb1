"""Per-phase device time of the sharded tensor search on ONE shard of an 8-way split (125k rows, 10k queries, k=100):
what is left besides the contraction when the shard gets small?"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lab_1806_vec_db_b200 as V
from lab_1806_vec_db_b200 import _lib as L
from bench import synth, load_fixtures, DIM
n = int(os.environ.get("N", 125_000)); nq = 10_000; k = 100
dev = torch.device("cuda:0")
b1000, t1000 = load_fixtures()
base = synth(b1000, 0, n, 42, dev)
q = synth(t1000, 0, nq, 43, dev)
vs = V.DeviceVecSet.from_device(base.data_ptr(), n, DIM, DIM, np.float32, "l2sqr", keepalive=base)
lib = L.lib()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
nn, ns, mn = C.c_uint64(0), C.c_uint32(0), C.c_float(0)
me = C.c_float(0)
L.check(lib.vdb_tq_info(vs._h, C.byref(nn), C.byref(ns), C.byref(mn), C.byref(me)))
j0 = int(os.environ.get("J0", 0)) or int(lib.vdb_tq_j0(k, ns.value, nn.value)); j = int(os.environ.get("J", 0)) or j0  # J0=2 on a 125k shard ~ the global threshold of an 8-way split
print("n", nn.value, "sample", ns.value, "j0", j0)
names = ["begin", "sample", "tau", "filter", "check", "decode"]
def once(record):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    tq = C.c_void_p()
    ev[0].record()
    L.check(lib.vdb_tq_begin_dev(vs._h, C.c_void_p(q.data_ptr()), nq, st, C.byref(tq))); ev[1].record()
    jkeys = torch.empty((nq, j), dtype=torch.int64, device=dev)
    L.check(lib.vdb_tq_sample_dev(tq, j, C.c_void_p(jkeys.data_ptr()))); ev[2].record()
    tau = torch.empty((nq,), dtype=torch.float32, device=dev)
    L.check(lib.vdb_tq_tau_dev(tq, C.c_void_p(jkeys.data_ptr()), 1, j, j0, C.c_void_p(tau.data_ptr()))); ev[3].record()
    keys = torch.empty((nq, k), dtype=torch.int64, device=dev); ovf = torch.empty((nq,), dtype=torch.int32, device=dev)
    L.check(lib.vdb_tq_filter_dev(tq, k, C.c_void_p(tau.data_ptr()), C.c_void_p(keys.data_ptr()), C.c_void_p(ovf.data_ptr()))); ev[4].record()
    redo = torch.empty((nq,), dtype=torch.int32, device=dev); nredo = torch.zeros((1,), dtype=torch.int32, device=dev)
    L.check(lib.vdb_tq_check_dev(tq, C.c_void_p(keys.data_ptr()), k, nn.value, C.c_void_p(tau.data_ptr()), C.c_void_p(ovf.data_ptr()),
                                 C.c_void_p(redo.data_ptr()), C.c_void_p(nredo.data_ptr())))
    nr = int(nredo.item()); ev[5].record()
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev); dd = torch.empty((nq, k), dtype=torch.float32, device=dev)
    cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
    L.check(lib.vdb_decode_keys_dev(C.c_void_p(keys.data_ptr()), nq, k, C.c_void_p(ids.data_ptr()), C.c_void_p(dd.data_ptr()),
                                    C.c_void_p(cnt.data_ptr()), st)); ev[6].record()
    lib.vdb_tq_end(tq)
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(len(names))], nr
for parts in [int(v) for v in os.environ.get("PARTS", "0").split(",")]:
    os.environ["VDB_GEMM_PARTS"] = str(parts)
    for _ in range(3): once(False)
    L.check(lib.vdb_prof_reset()); L.check(lib.vdb_prof_enable(1))
    acc = np.zeros(len(names)); R = 5
    for _ in range(R):
        t, nr = once(True); acc += np.array(t)
    L.check(lib.vdb_prof_enable(0))
    print("parts", parts, "phase ms:", {nm: round(float(v) / R, 3) for nm, v in zip(names, acc)}, "total", round(float(acc.sum()) / R, 3), "redo", nr)
    out = {}
    for name in (b"flat_gemm", b"rerank", b"merge"):
        t, c = C.c_double(0), C.c_uint64(0)
        L.check(lib.vdb_prof_read(name, C.byref(t), C.byref(c)))
        out[name.decode()] = (round(t.value / R, 3), c.value // R)
    print("   kernels (ms, launches):", out)
