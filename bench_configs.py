"""The other single-GPU configs of BASELINE.json measured like the headline (bench.py imports this at N = 1):

  configs[2]  IVF   nlist = 128, k-means on a 100k-row sample, nprobe in {8, 16, 24}      (config/bench_10000_ivf.toml:7-16)
  configs[3]  PQ    m = 240 x 4 bits, k-means on 10k rows, Flat+PQ scan, ef in {240, 420, 600}  (config/bench_pq_240_hnsw.toml:7-23)
  configs[4]  HNSW  M = 16, ef_construction = 200, ef in {120, 200, 360}; HNSW+PQ ef in {240, 420, 600}  (config/bench_hnsw.toml:7-14)
  u8 rows     Flat L2Sqr on the byte-quantised set (the reference's second scalar type, src/scalar.rs:117-119)
  cosine      Flat cosine on the headline set: single query and the 10 000-query batch (src/distance/mod.rs:60-69)

1000 queries, k = 10, recall@10 against the exact Flat result (examples/bench.rs protocol, src/bin/gen_gnd.rs ground
truth). Every search row carries
  qps / ms_per_batch   device-resident, CUDA events on the launching stream
  e2e                  the same call through the host-pointer C ABI (H2D of the queries + D2H of the results inside)
  roofline             algorithmic bytes or FLOPs of SURVEY.md section 8d / DESIGN.md divided by the CUDA-event time and
                       by the measured peak of MEASURED_PEAKS.json
  cpu_baseline         the oracle (reference semantics restated in C++) on a bounded query sample with all host
                       threads IN THE SAME RUN: its QPS, ITS recall on that sample, and the rate at which the GPU ids
                       equal the oracle's on the same centroids / codebooks / graph
Two synthetic sets: "flat" (bench.synth: 1000 near-equidistant copies per prototype - PQ / HNSW can only rank them by
chance) and "clustered" (bench.synth_clustered: low-rank clouds - the set on which QPS can be quoted at the reference's
recall range 0.85-0.95, data/t_bench.toml:4-23, data/t_bench_pq.toml:4-23).
"""
import ctypes as C
import os
import time

import numpy as np

DIM = 960


def _recall(ids, gt, k):
    return float(np.mean([len(set(a) & set(b)) / k for a, b in zip(np.asarray(ids).tolist(), np.asarray(gt).tolist())]))


class Legs:
    def __init__(self, V, L, lib, dev, peaks, name, base, q_dev, cpu_queries=16):
        import torch
        self.torch, self.V, self.L, self.lib, self.dev, self.peaks, self.name = torch, V, L, lib, dev, peaks, name
        self.base, self.n = base, base.shape[0]
        self.nq, self.k = min(1000, q_dev.shape[0]), 10
        self.q = q_dev[:self.nq].contiguous()
        self.q_host = self.q.cpu().numpy()
        self.vs = V.DeviceVecSet.from_device(base.data_ptr(), self.n, DIM, DIM, np.float32, "l2sqr", keepalive=base)
        self.flat = V.FlatIndex(self.vs)
        self.gt = self.flat.knn_batch(self.q_host, self.k)[0].astype(np.int64)      # exact ground truth (gen_gnd.rs)
        self.cores = os.cpu_count() or 1
        self.nc = min(cpu_queries, self.nq)
        self.st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        self.hbm = peaks.get("hbm_gbs") or 6650.0
        self.bf16 = peaks.get("bf16_tflops_sustained") or 1403.0
        self.peak_src = "MEASURED_PEAKS.json" if peaks.get("hbm_gbs") else "fallback (B200_PROFILING.md)"
        self.rng = np.random.default_rng(42)
        self._base_host = None
        self.ids = torch.empty((self.nq, self.k), dtype=torch.int64, device=dev)
        self.dd = torch.empty((self.nq, self.k), dtype=torch.float32, device=dev)
        self.cnt = torch.empty((self.nq,), dtype=torch.int32, device=dev)

    # ---- helpers ----
    def base_host(self):
        if self._base_host is None:
            self._base_host = self.base.cpu().numpy()
        return self._base_host

    def sample_rows(self, m):
        sel = self.torch.as_tensor(self.rng.permutation(self.n)[:min(m, self.n)], device=self.dev)
        return np.ascontiguousarray(self.base.index_select(0, sel).cpu().numpy())

    def timed(self, fn, reps=3):
        t = self.torch
        fn()
        t.cuda.synchronize()
        a0, a1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(reps):
            fn()
        a1.record()
        t.cuda.synchronize()
        return a0.elapsed_time(a1) / reps

    def wall(self, fn, reps=3):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            out = fn()
        return (time.perf_counter() - t0) / reps, out

    def out_ptrs(self):
        return (C.c_void_p(self.ids.data_ptr()), C.c_void_p(self.dd.data_ptr()), C.c_void_p(self.cnt.data_ptr()))

    def row(self, param, ms, e2e_s, e2e_api, roofline, cpu_ids, cpu_s, what):
        """One search row; cpu_ids = the oracle's ids for the first nc queries (same trained artefacts)."""
        got = self.ids.cpu().numpy()
        r = dict(param)
        r.update({"qps": self.nq / ms * 1e3, "ms_per_batch": ms, "recall@10": _recall(got, self.gt, self.k),
                  "e2e": {"value": self.nq / e2e_s, "unit": "queries/s", "ms_per_batch": e2e_s * 1e3, "api": e2e_api,
                          "h2d_bytes": self.nq * DIM * 4, "d2h_bytes": self.nq * self.k * 12 + self.nq * 4},
                  "roofline": roofline})
        if cpu_ids is not None:
            r["cpu_baseline"] = {"value": self.nc / cpu_s, "unit": "queries/s", "cores": self.cores, "kind": "port",
                                 "sample": f"{self.nc} of the {self.nq} queries, {what}",
                                 "recall@10": _recall(cpu_ids, self.gt[:self.nc], self.k),
                                 "gpu_recall@10_same_sample": _recall(got[:self.nc], self.gt[:self.nc], self.k),
                                 "gpu_ids_equal_oracle_rate": float((got[:self.nc] == np.asarray(cpu_ids).astype(np.int64)).mean())}
        return r

    # ---- configs[2]: IVF ----
    def ivf(self, nprobes=(8, 16, 24)):
        import oracle as O
        V, L, lib = self.V, self.L, self.lib
        t0 = time.perf_counter()
        km = V.KMeans.from_vec_set(self.sample_rows(100_000), V.KMeansConfig(128, 20, 1e-6, "l2sqr"), self.rng)
        t_train = time.perf_counter() - t0
        t0 = time.perf_counter()
        ivf = V.IVFIndex(self.vs, km.centroids)
        t_assign = time.perf_counter() - t0
        off, mem = O.ivf_lists(ivf.assignment, 128)
        sizes = np.diff(np.asarray(off).astype(np.int64))
        rows = []
        for nprobe in nprobes:
            ms = self.timed(lambda: L.check(lib.vdb_ivf_knn_dev(self.vs._h, ivf._h, C.c_void_p(self.q.data_ptr()), self.nq, self.k,
                                                                nprobe, *self.out_ptrs(), self.st)))
            e2e_s, _ = self.wall(lambda: ivf.knn_with_ef_batch(self.q_host, self.k, nprobe))
            bh = self.base_host()   # the host copy of the rows is made outside the timed call
            t0 = time.perf_counter()
            oi, _, _ = O.ivf_knn(bh, km.centroids, off, mem, self.q_host[:self.nc], self.k, nprobe, "l2sqr",
                                 nthreads=self.cores)
            cpu_s = time.perf_counter() - t0
            # a 1000-query batch probes every list: each list's rows are streamed once (FP16 operand copy in list order)
            hbm_bytes = self.n * DIM * 2 + self.nq * DIM * 2
            flops = 2.0 * DIM * self.nq * nprobe * float(sizes.mean())
            roof = {"bound": "hbm", "achieved": hbm_bytes / (ms * 1e-3) / 1e9, "peak": self.hbm, "unit": "GB/s",
                    "frac": hbm_bytes / (ms * 1e-3) / 1e9 / self.hbm, "peak_source": self.peak_src,
                    "algorithmic_bytes_per_batch": hbm_bytes,
                    "basis": "every list is probed by some query of the batch, so the probe scan reads each row once "
                             "(2-byte operands) - SURVEY.md 8d's per-query figure summed over the batch would count a list "
                             "once per probing query",
                    "tensor_tflops": flops / (ms * 1e-3) / 1e12, "tensor_frac_of_bf16_sustained": flops / (ms * 1e-3) / 1e12 / self.bf16}
            rows.append(self.row({"nprobe": nprobe}, ms, e2e_s, "vdb_ivf_knn (host pointers)", roof, oi, cpu_s,
                                 "oracle ivf_knn (find_n_nearest + probe scan), thread pool over queries"))
        return {"reference_config": "config/bench_10000_ivf.toml:7-16 scaled x100", "nlist": 128,
                "kmeans_rows": min(100_000, self.n), "kmeans_iters": int(km.iterations), "train_s": t_train,
                "assign_and_lists_s": t_assign, "list_min_max": [int(sizes.min()), int(sizes.max())], "search": rows}

    # ---- configs[3]: PQ table + Flat ADC scan + exact rerank ----
    def pq_table(self):
        from lab_1806_vec_db_b200.index import train_codebooks
        V = self.V
        cfg = V.PQConfig(4, 240, "l2sqr", min(10_000, self.n), 20, 1e-6)
        t0 = time.perf_counter()
        train_dev = V.DeviceVecSet(self.sample_rows(10_000), "l2sqr")
        books = train_codebooks(train_dev, cfg, self.rng)
        train_dev.close()
        t_train = time.perf_counter() - t0
        t0 = time.perf_counter()
        pq = V.PQTable(self.vs, cfg, books)
        return pq, books, t_train, time.perf_counter() - t0

    def pq(self, pq, books, t_train, t_encode, efs=(240, 420, 600)):
        import oracle as O
        L, lib = self.L, self.lib
        rows = []
        for ef in efs:
            ms = self.timed(lambda: L.check(lib.vdb_pq_knn_dev(self.vs._h, pq._h, C.c_void_p(self.q.data_ptr()), self.nq, self.k,
                                                               ef, *self.out_ptrs(), self.st)))
            e2e_s, _ = self.wall(lambda: self.flat.knn_pq_batch(self.q_host, self.k, ef, pq))
            bh = self.base_host()
            t0 = time.perf_counter()
            oi, _, _ = O.flat_knn_pq(bh, pq.encoded_vec_set, books, 240, 4, self.q_host[:self.nc], self.k, ef,
                                     "l2sqr", nthreads=self.cores)
            cpu_s = time.perf_counter() - t0
            # m = dim / 4: the ADC filter is a contraction over rows decoded on the fly, K = dim; the one-hot
            # form of round 1 (K = 16 m) did 4x the FLOP for the same scan
            flops = 2.0 * self.nq * self.n * DIM
            roof = {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": self.bf16, "unit": "TFLOP/s",
                    "frac": flops / (ms * 1e-3) / 1e12 / self.bf16, "peak_source": self.peak_src + " bf16_tflops_sustained",
                    "flop_per_batch": flops, "code_bytes_per_pass": self.n * 120,
                    "basis": "2 nq n dim FLOP (fp16 operands: decoded centroid values x queries) divided by the time of the WHOLE "
                             "call (LUT, sample pass, contraction, exact re-evaluation, merges, rerank); the contraction kernel "
                             "alone is 1.70 of the 2.63 ms (profiles/r02_pq_dec_launches.md)"}
            rows.append(self.row({"ef": ef}, ms, e2e_s, "vdb_pq_knn (host pointers)", roof, oi, cpu_s,
                                 "oracle knn_pq (ADC scan of all codes + pq_resort), thread pool over queries"))
        return {"reference_config": "config/bench_pq_240_hnsw.toml:16-23 (Flat+PQ scan variant)", "m": 240, "n_bits": 4,
                "kmeans_rows": min(10_000, self.n), "train_s_incl_upload": t_train, "encode_s_incl_code_download": t_encode,
                "search": rows}

    # ---- configs[4]: HNSW (+PQ) ----
    def hnsw(self, pq=None, books=None, efs=(120, 200, 360), pq_efs=(240, 420, 600)):
        from oracle.oracle_py import HnswOracle
        V, L, lib = self.V, self.L, self.lib
        t0 = time.perf_counter()
        hn = V.HNSWIndex(self.vs, V.HNSWConfig(0, 200, 16), rng=np.random.default_rng(42))
        t_build = time.perf_counter() - t0
        links0, len0 = hn.level0_links()
        levels, ulinks, ulen = hn.upper_links()
        ep, el = hn.enter_point
        cpu = HnswOracle.from_graph(self.base_host(), "l2sqr", 16, 200, levels, links0, len0, ulinks, ulen, ep, el)
        out = {"reference_config": "config/bench_hnsw.toml:7-14", "M": 16, "ef_construction": 200, "build_s": t_build,
               "inserts_per_s": self.n / t_build, "level0_degree_mean": float(len0.mean()),
               "cpu_baseline_graph": "the oracle's knn_with_ef / knn_pq walk the graph the GPU built (a CPU build of 1M "
                                     "rows takes ~0.5 h on one thread): search parity on an identical graph",
               "search": [], "search_pq": []}

        def evals_of(fn):
            """distance evaluations of ONE call (the library counts the rows its search_on_level loops gather)"""
            n_ev = C.c_uint64(0)
            L.check(lib.vdb_hnsw_evals(hn._h, C.byref(n_ev), 1))
            fn()
            self.torch.cuda.synchronize()
            L.check(lib.vdb_hnsw_evals(hn._h, C.byref(n_ev), 1))
            return int(n_ev.value)

        def roof(ms, evals, unit_bytes, what):
            gbs = evals * unit_bytes / (ms * 1e-3) / 1e9
            return {"bound": "hbm (dependent gathers: latency, not bandwidth)", "achieved": gbs, "peak": self.hbm, "unit": "GB/s",
                    "frac": gbs / self.hbm, "peak_source": self.peak_src, "evaluations_per_query": evals / self.nq,
                    "algorithmic_bytes_per_batch": evals * unit_bytes,
                    "basis": f"SURVEY.md 8d: evaluations x {what}, evaluations counted by the search kernel (vdb_hnsw_evals)",
                    "ms_per_batch": ms}
        for ef in efs:
            call = lambda: L.check(lib.vdb_hnsw_knn_dev(self.vs._h, hn._h, C.c_void_p(self.q.data_ptr()), self.nq, self.k,  # noqa: E731
                                                        ef, *self.out_ptrs(), self.st))
            evals = evals_of(call)
            ms = self.timed(call)
            e2e_s, _ = self.wall(lambda: hn.knn_with_ef_batch(self.q_host, self.k, ef))
            t0 = time.perf_counter()
            oi, _, _ = cpu.knn(self.q_host[:self.nc], self.k, ef, nthreads=self.cores)
            cpu_s = time.perf_counter() - t0
            out["search"].append(self.row({"ef": ef}, ms, e2e_s, "vdb_hnsw_knn (host pointers)",
                                          roof(ms, evals, DIM * 4 + 4, "(dim 4 + 4) bytes per gathered row"), oi, cpu_s,
                                          "oracle knn_with_ef on the same graph, thread pool over queries"))
        if pq is not None:
            for ef in pq_efs:
                call = lambda: L.check(lib.vdb_hnsw_knn_pq_dev(self.vs._h, hn._h, pq._h, C.c_void_p(self.q.data_ptr()),  # noqa: E731
                                                               self.nq, self.k, ef, *self.out_ptrs(), self.st))
                evals = evals_of(call)
                ms = self.timed(call)
                e2e_s, _ = self.wall(lambda: hn.knn_pq_batch(self.q_host, self.k, ef, pq))
                t0 = time.perf_counter()
                oi, _, _ = cpu.knn_pq(self.q_host[:self.nc], self.k, ef, pq.encoded_vec_set, books, 240, 4, nthreads=self.cores)
                cpu_s = time.perf_counter() - t0
                out["search_pq"].append(self.row({"ef": ef}, ms, e2e_s, "vdb_hnsw_knn_pq (host pointers)",
                                                 roof(ms, evals, 120 + 4, "(120-byte code + 4) bytes per walked node; the exact "
                                                      "rerank of max(ef, k) rows is not counted"), oi, cpu_s,
                                                 "oracle knn_pq on the same graph and codes, thread pool over queries"))
        return out


def u8_leg(V, L, lib, dev, peaks, base_f32, q_f32):
    """Flat L2Sqr over u8 rows (the set scaled to 0..255): the trait's single-query call and a 1000-query batch."""
    import torch
    import oracle as O
    n = base_f32.shape[0]
    base = (base_f32 * 255.0).round_().clamp_(0, 255).to(torch.uint8)
    q = (q_f32[:1000] * 255.0).round_().clamp_(0, 255).to(torch.uint8).contiguous()
    vs = V.DeviceVecSet.from_device(base.data_ptr(), n, DIM, DIM, np.uint8, "l2sqr", keepalive=base)
    flat = V.FlatIndex(vs)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    hbm = peaks.get("hbm_gbs") or 6650.0
    out = {"rows": "round(255 x) of the headline set, u8", "cases": []}
    q_host = q.cpu().numpy()
    cores = os.cpu_count() or 1
    for nq, k in ((1, 10), (1000, 10)):
        qq = q[:nq].contiguous()
        ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
        dd = torch.empty((nq, k), dtype=torch.float32, device=dev)
        cnt = torch.empty((nq,), dtype=torch.int32, device=dev)

        def run():
            L.check(lib.vdb_flat_knn_dev(vs._h, C.c_void_p(qq.data_ptr()), nq, k, C.c_void_p(ids.data_ptr()),
                                         C.c_void_p(dd.data_ptr()), C.c_void_p(cnt.data_ptr()), st))
        for _ in range(5):
            run()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20 if nq == 1 else 5
        a0.record()
        for _ in range(reps):
            run()
        a1.record()
        torch.cuda.synchronize()
        ms = a0.elapsed_time(a1) / reps
        flat.knn_batch(q_host[:nq], k)
        t0 = time.perf_counter()
        for _ in range(reps):
            flat.knn_batch(q_host[:nq], k)
        e2e_s = (time.perf_counter() - t0) / reps
        case = {"nq": nq, "k": k, "qps": nq / ms * 1e3, "ms_per_call": ms,
                "e2e": {"value": nq / e2e_s, "unit": "queries/s", "api": "vdb_flat_knn (host pointers)"}}
        if nq == 1:
            case["roofline"] = {"bound": "hbm", "achieved": n * DIM / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                                "frac": n * DIM / (ms * 1e-3) / 1e9 / hbm, "algorithmic_bytes_per_call": n * DIM,
                                "basis": "whole call (query preparation + scan + merge + decode), n dim 1 bytes"}
            nc = 4
            t0 = time.perf_counter()
            oi, od, _ = O.flat_knn(base.cpu().numpy(), q_host[:nc], k, "l2sqr", nthreads=cores)
            cpu_s = time.perf_counter() - t0
            gi = flat.knn_batch(q_host[:nc], k)
            case["cpu_baseline"] = {"value": nc / cpu_s, "unit": "queries/s", "cores": cores, "kind": "port",
                                    "sample": f"{nc} queries x {n} u8 rows",
                                    "gpu_ids_equal_oracle_rate": float((gi[0].astype(np.int64) == oi.astype(np.int64)).mean()),
                                    "gpu_dist_bits_equal_oracle_rate": float((gi[1].view(np.uint32) == od.view(np.uint32)).mean())}
        out["cases"].append(case)
    return out


def cosine_leg(V, L, lib, dev, peaks, base, q_f32):
    """The headline batch (10 000 queries, k = 100) and the single-query call with the reference's other metric: cosine
    distance (src/distance/mod.rs:60-69) through the same kernels, checked against the oracle on a query sample."""
    import torch
    import oracle as O
    n = base.shape[0]
    vs = V.DeviceVecSet.from_device(base.data_ptr(), n, DIM, DIM, np.float32, "cosine", keepalive=base)
    flat = V.FlatIndex(vs)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    hbm = peaks.get("hbm_gbs") or 6650.0
    cores = os.cpu_count() or 1
    out = {"metric": "cosine", "cases": []}
    for nq, k, reps in ((1, 10, 30), (q_f32.shape[0], 100, 3)):
        qq = q_f32[:nq].contiguous()
        ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
        dd = torch.empty((nq, k), dtype=torch.float32, device=dev)
        cnt = torch.empty((nq,), dtype=torch.int32, device=dev)

        def run():
            L.check(lib.vdb_flat_knn_dev(vs._h, C.c_void_p(qq.data_ptr()), nq, k, C.c_void_p(ids.data_ptr()),
                                         C.c_void_p(dd.data_ptr()), C.c_void_p(cnt.data_ptr()), st))
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(reps):
            run()
        a1.record()
        torch.cuda.synchronize()
        ms = a0.elapsed_time(a1) / reps
        case = {"nq": nq, "k": k, "qps": nq / ms * 1e3, "ms_per_call": ms}
        if nq == 1:
            case["roofline"] = {"bound": "hbm", "achieved": n * DIM * 4 / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                                "frac": n * DIM * 4 / (ms * 1e-3) / 1e9 / hbm, "basis": "whole call, n dim 4 bytes"}
        else:
            nc = 8
            q_host = qq[:nc].cpu().numpy()
            t0 = time.perf_counter()
            oi, od, _ = O.flat_knn(base.cpu().numpy(), q_host, k, "cosine", nthreads=cores)
            cpu_s = time.perf_counter() - t0
            gi, gd = ids[:nc].cpu().numpy(), dd[:nc].cpu().numpy()
            # an id mismatch must be a tie: the ORACLE's distance of the returned row within 1e-5 (absolute: 1 - cos
            # cancels) of the oracle's distance at that rank
            ties_ok = True
            for qi, j in zip(*np.nonzero(gi != oi.astype(np.int64))):
                d = O.distance(q_host[qi], base[int(gi[qi, j])].cpu().numpy(), "cosine")
                ties_ok &= abs(d - od[qi, j]) <= 1e-5
            case["cpu_baseline"] = {"value": nc / cpu_s, "unit": "queries/s", "cores": cores, "kind": "port",
                                    "sample": f"{nc} of the {nq} queries",
                                    "gpu_ids_equal_oracle_rate": float((gi == oi.astype(np.int64)).mean()),
                                    "id_mismatches_are_ties_within_1e-5_abs": bool(ties_ok),
                                    "gpu_max_abs_dist_err": float(np.abs(gd - od).max())}
        out["cases"].append(case)
    return out


def run_all(V, L, lib, dev, peaks, base_flat, q_flat, synth_clustered, proto, n):
    """Both sets. The flat set is the headline's (already resident); the clustered one is generated here."""
    import torch
    out = {"protocol": "1000 queries, k = 10, recall@10 vs the exact Flat result (examples/bench.rs / src/bin/gen_gnd.rs)"}
    legs = Legs(V, L, lib, dev, peaks, "flat", base_flat, q_flat)
    flat_set = {"data": "bench.synth: prototype + 0.02 N(0,1): 1000 near-equidistant copies per prototype (PQ / HNSW rank them "
                        "by chance; recall equals the oracle's, see cpu_baseline.recall@10)"}
    flat_set["ivf"] = legs.ivf()
    pq, books, t_train, t_encode = legs.pq_table()
    flat_set["pq"] = legs.pq(pq, books, t_train, t_encode)
    out["flat_set"] = flat_set
    try:
        out["u8"] = u8_leg(V, L, lib, dev, peaks, base_flat, q_flat)
    except Exception as e:  # noqa: BLE001
        out["u8"] = {"error": repr(e)}
    try:
        out["cosine"] = cosine_leg(V, L, lib, dev, peaks, base_flat, q_flat)
    except Exception as e:  # noqa: BLE001
        out["cosine"] = {"error": repr(e)}
    del legs, pq
    torch.cuda.empty_cache()
    base2 = synth_clustered(proto, 0, n, 42, dev)
    q2 = synth_clustered(proto, 0, 1000, 43, dev)
    legs = Legs(V, L, lib, dev, peaks, "clustered", base2, q2)
    cl = {"data": "bench.synth_clustered(latent=32, spread=0.010, noise=0.018): prototype + low-rank cloud + isotropic noise; "
                  "queries are fresh members of the same mixture (scripts/probe_recall.py calibrated it to the reference's "
                  "recall range)"}
    cl["ivf"] = legs.ivf()
    pq, books, t_train, t_encode = legs.pq_table()
    cl["pq"] = legs.pq(pq, books, t_train, t_encode)
    cl["hnsw"] = legs.hnsw(pq, books)
    out["clustered_set"] = cl
    return out


def sharded_legs(V, vs, base_host, q_host, gt10, n_gpus):
    """configs[2] / configs[3] on the ROW-SHARDED set at N > 1 GPUs, through the host-pointer C calls a Rust host makes
    (vdb_ivf_create / vdb_ivf_knn, vdb_pq_create / vdb_pq_knn on the sharded handle: csrc/multi.cu). 1000 queries, k = 10;
    wall-clock QPS including the H2D of the queries and the D2H of the results, recall@10 against the exact Flat result
    of the headline leg, and the oracle on a query sample in the same run (QPS, its recall, id equality)."""
    import oracle as O
    from lab_1806_vec_db_b200.index import train_codebooks
    nq, k = min(1000, q_host.shape[0]), 10
    q = np.ascontiguousarray(q_host[:nq])
    gt = np.asarray(gt10)[:nq, :k].astype(np.int64)
    cores = os.cpu_count() or 1
    nc = min(16, nq)
    rng = np.random.default_rng(42)
    n = base_host.shape[0]
    out = {"n_gpus": n_gpus, "nq": nq, "k": k, "api": "host pointers on the row-sharded handle (one C call per batch)"}

    def wall(fn, reps=3):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = fn()
        return (time.perf_counter() - t0) / reps, r

    def row(param, secs, res, oi, cpu_s, what):
        got = res[0].astype(np.int64)
        return dict(param, **{"e2e_qps": nq / secs, "ms_per_batch": secs * 1e3, "recall@10": _recall(got, gt, k),
                              "cpu_baseline": {"value": nc / cpu_s, "unit": "queries/s", "cores": cores, "kind": "port",
                                               "sample": f"{nc} of the {nq} queries, {what}",
                                               "recall@10": _recall(oi, gt[:nc], k),
                                               "gpu_ids_equal_oracle_rate": float((got[:nc] == np.asarray(oi).astype(np.int64)).mean())}})

    # ---- IVF ----
    km = V.KMeans.from_vec_set(np.ascontiguousarray(base_host[rng.permutation(n)[:100_000]]),
                               V.KMeansConfig(128, 20, 1e-6, "l2sqr"), rng)
    t0 = time.perf_counter()
    ivf = V.IVFIndex(vs, km.centroids)
    t_build = time.perf_counter() - t0
    off, mem = O.ivf_lists(ivf.assignment, 128)
    rows = []
    for nprobe in (8, 24):
        secs, res = wall(lambda: ivf.knn_with_ef_batch(q, k, nprobe))
        t0 = time.perf_counter()
        oi, _, _ = O.ivf_knn(base_host, km.centroids, off, mem, q[:nc], k, nprobe, "l2sqr", nthreads=cores)
        rows.append(row({"nprobe": nprobe}, secs, res, oi, time.perf_counter() - t0, "oracle ivf_knn"))
    out["ivf"] = {"nlist": 128, "assign_and_lists_s": t_build, "search": rows}
    ivf.close()
    # ---- PQ ----
    cfg = V.PQConfig(4, 240, "l2sqr", min(10_000, n), 20, 1e-6)
    train_dev = V.DeviceVecSet(np.ascontiguousarray(base_host[rng.permutation(n)[:10_000]]), "l2sqr")
    books = train_codebooks(train_dev, cfg, rng)
    train_dev.close()
    t0 = time.perf_counter()
    pq = V.PQTable(vs, cfg, books)
    t_enc = time.perf_counter() - t0
    flat = V.FlatIndex(vs)
    rows = []
    for ef in (240, 600):
        secs, res = wall(lambda: flat.knn_pq_batch(q, k, ef, pq))
        t0 = time.perf_counter()
        oi, _, _ = O.flat_knn_pq(base_host, pq.encoded_vec_set, books, 240, 4, q[:nc], k, ef, "l2sqr", nthreads=cores)
        rows.append(row({"ef": ef}, secs, res, oi, time.perf_counter() - t0, "oracle knn_pq"))
    out["pq"] = {"m": 240, "n_bits": 4, "encode_s_incl_code_download": t_enc, "search": rows}
    return out
