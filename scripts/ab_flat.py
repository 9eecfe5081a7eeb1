#!/usr/bin/env python
"""A/B timing of the batched Flat search inside ONE process (boxes differ by several per cent under the power cap, so
variants are compared interleaved on the same GPU). Variants are environment settings the library reads per call.

    python scripts/ab_flat.py "VDB_GEMM_RARE_PER_SCORE=1" "VDB_GEMM_PARTS=3"     (the empty variant is always included)
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import lab_1806_vec_db_b200 as V
    from lab_1806_vec_db_b200 import _lib as L
    n, nq, k = int(os.environ.get("N", 1_000_000)), int(os.environ.get("NQ", 10_000)), int(os.environ.get("K", 100))
    dev = torch.device("cuda:0")
    lib = L.lib()
    L.check(lib.vdb_set_device(0))
    b1000, t1000 = bench.load_fixtures()
    base = bench.synth(b1000, 0, n, 42, dev)
    q = bench.synth(t1000, 0, nq, 43, dev)
    vs = V.DeviceVecSet.from_device(base.data_ptr(), n, bench.DIM, bench.DIM, np.float32, "l2sqr", keepalive=base)
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    dd = torch.empty((nq, k), dtype=torch.float32, device=dev)
    cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def call():
        L.check(lib.vdb_flat_knn_dev(vs._h, C.c_void_p(q.data_ptr()), nq, k, C.c_void_p(ids.data_ptr()), C.c_void_p(dd.data_ptr()),
                                     C.c_void_p(cnt.data_ptr()), st))
    variants = [""] + sys.argv[1:]
    names = ("flat_gemm", "flat_gemm_sample", "rerank", "merge", "flat_scan")
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    ref = None
    for rnd in range(int(os.environ.get("ROUNDS", 3))):
        for v in variants:
            sets = [kv.split("=", 1) for kv in v.split(",") if kv]
            for key, val in sets:
                os.environ[key] = val
            call()
            torch.cuda.synchronize()
            L.check(lib.vdb_prof_reset()); L.check(lib.vdb_prof_enable(1))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = int(os.environ.get("REPS", 5))
            e0.record()
            for _ in range(reps):
                call()
            e1.record()
            torch.cuda.synchronize()
            L.check(lib.vdb_prof_enable(0))
            prof = {}
            for nm in names:
                t, c = C.c_double(0), C.c_uint64(0)
                L.check(lib.vdb_prof_read(nm.encode(), C.byref(t), C.byref(c)))
                prof[nm] = round(t.value / reps, 3)
            same = None
            if ref is None:
                ref = ids.clone()
            else:
                same = bool((ids == ref).all())
            print(f"round {rnd} [{v or 'default':40s}] {e0.elapsed_time(e1) / reps:8.3f} ms/step  {prof}  ids==first:{same}", flush=True)
            for key, _ in sets:
                del os.environ[key]


if __name__ == "__main__":
    main()
