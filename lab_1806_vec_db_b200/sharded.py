"""Row-sharded Flat search: one process per GPU, database rows split into contiguous blocks,
per-GPU top-k merged after one NCCL all-gather of [nq, k] packed keys.

torch is used here only as plumbing (device buffers, streams, torch.distributed); every compute
step is a kernel of libvdb_b200.so.
"""
import ctypes as C
import os

import numpy as np

from . import _lib as L

KEY_NONE = np.uint64(0xFFFFFFFFFFFFFFFF)


def shard_bounds(n, world, rank):
    """Contiguous row block [lo, hi) of `rank`; keeps 'lower id wins ties' identical to the unsharded scan."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_keys(dist, ids):
    """Host mirror of make_key (csrc/common.cuh): u64 keys whose integer order is (distance, id)."""
    d = np.asarray(dist, np.float32) + np.float32(0.0)
    b = d.view(np.uint32).copy()
    b[np.isnan(d)] = 0x7FC00000
    neg = (b & 0x80000000) != 0
    o = np.where(neg, ~b, b | np.uint32(0x80000000)).astype(np.uint64)
    return (o << np.uint64(32)) | np.asarray(ids, np.uint64)


def unpack_keys(keys):
    keys = np.asarray(keys, np.uint64)
    o = (keys >> np.uint64(32)).astype(np.uint32)
    b = np.where((o & 0x80000000) != 0, o & np.uint32(0x7FFFFFFF), ~o).astype(np.uint32)
    return b.view(np.float32), (keys & np.uint64(0xFFFFFFFF)).astype(np.uint64)


class ShardedFlatIndex:
    """Flat index over this rank's row block; search results are global and identical on all ranks.

    Large L2Sqr/f32 batches use the tensor-core phases of the library with GLOBAL thresholds: the shards'
    sample scores are all-gathered so that every shard filters against the same tau_q and reranks only its share
    of the ~k candidates (instead of its own ~750 per query); small batches use the exact streaming scan.
    Collectives per batch: all-gather of [nq, j0] sample keys, all-gather of [nq, k] result keys, all-reduce of
    the per-query overflow flags (tensor path only)."""

    TENSOR_MIN_NQ = 12
    QUERY_CHUNK = 16384
    PIPELINE_CHUNKS = int(os.environ.get("VDB_GEMM_CHUNKS", "1"))  # > 1: two-stream chunk pipeline (measured: no gain)

    def __init__(self, vec_set, rank=0, world=1):
        self.vec_set = vec_set
        self.rank, self.world = rank, world
        self._tensor = None  # (n_total, sample_total, mean_norm) once known; False if unsupported
        self._streams = None
        self.last_fallbacks = 0
        self.phase_timing = bool(int(os.environ.get("VDB_PHASE_TIMING", "0")))
        self.phase_ms = {}

    # ---- helpers -------------------------------------------------------------------------------------------
    def _tensor_info(self, dev):
        if self._tensor is None:
            import torch
            lib = L.lib()
            n, ns, mn, me = C.c_uint64(0), C.c_uint32(0), C.c_float(0), C.c_float(0)
            ok = (self.vec_set.dtype in (np.float32, np.uint8) and
                  lib.vdb_tq_info(self.vec_set._h, C.byref(n), C.byref(ns), C.byref(mn), C.byref(me)) == L.OK)
            t = torch.tensor([float(n.value), float(ns.value), 1.0 if ok else 0.0], dtype=torch.float64, device=dev)
            m = torch.tensor([mn.value, me.value], dtype=torch.float32, device=dev)
            if self.world > 1:
                import torch.distributed as dist
                okmin = t[2:3].clone()
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                dist.all_reduce(okmin, op=dist.ReduceOp.MIN)
                dist.all_reduce(m, op=dist.ReduceOp.MAX)
                ok = bool(okmin.item() > 0)
            self._tensor = (int(t[0].item()), int(t[1].item()), float(m[0].item()), float(m[1].item())) if ok else False
        return self._tensor

    def _gather(self, t):
        import torch
        if self.world == 1:
            return t.unsqueeze(0)
        import torch.distributed as dist
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous())
        return out

    def _scan_keys(self, q, k, st):
        import torch
        keys = torch.empty((q.shape[0], k), dtype=torch.int64, device=q.device)
        L.check(L.lib().vdb_flat_scan_keys_dev(self.vec_set._h, C.c_void_p(q.data_ptr()), q.shape[0], k,
                                               C.c_void_p(keys.data_ptr()), st))
        return keys

    def _merge_to_keys(self, allk, nq, k, st):
        import torch
        out = torch.empty((nq, k), dtype=torch.int64, device=allk.device)
        L.check(L.lib().vdb_merge_keys_to_keys_dev(C.c_void_p(allk.data_ptr()), allk.shape[0], nq, k,
                                                   C.c_void_p(out.data_ptr()), st))
        return out

    def _tensor_enqueue(self, q, k, info):
        """Enqueue one chunk's tensor-core phases on the CURRENT torch stream without any host synchronisation.
        Returns the state `_tensor_finish` needs (the global [nq, k] keys are complete unless the check flags a query)."""
        import torch
        lib = L.lib()
        n_total, ns_total, mean_norm, mean_ex = info
        nq, dev = q.shape[0], q.device
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        j0 = int(lib.vdb_tq_j0(k, ns_total, n_total))
        j = int(lib.vdb_tq_sample_j(j0, max(1, ns_total // self.world)))
        tq = C.c_void_p()
        marks = []

        def mark(name):  # optional per-phase device timing (VDB_PHASE_TIMING=1): events on the chunk's stream
            if self.phase_timing:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append((name, ev))

        mark("start")
        L.check(lib.vdb_tq_begin_dev(self.vec_set._h, C.c_void_p(q.data_ptr()), nq, st, C.byref(tq)))
        try:
            mark("begin")
            jkeys = torch.empty((nq, j), dtype=torch.int64, device=dev)
            L.check(lib.vdb_tq_sample_dev(tq, j, C.c_void_p(jkeys.data_ptr())))
            mark("sample")
            allj = self._gather(jkeys)
            mark("gather_sample")
            tau = torch.empty((nq,), dtype=torch.float32, device=dev)
            L.check(lib.vdb_tq_tau_dev(tq, C.c_void_p(allj.data_ptr()), self.world, j, min(j0, j * self.world),
                                       C.c_void_p(tau.data_ptr())))
            mark("tau")
            keys = torch.empty((nq, k), dtype=torch.int64, device=dev)
            ovf = torch.empty((nq,), dtype=torch.int32, device=dev)
            L.check(lib.vdb_tq_filter_dev(tq, k, C.c_void_p(tau.data_ptr()), C.c_void_p(keys.data_ptr()),
                                          C.c_void_p(ovf.data_ptr())))
            mark("filter_rerank")
            allk = self._gather(keys)
            if self.world > 1:
                import torch.distributed as dist
                dist.all_reduce(ovf, op=dist.ReduceOp.MAX)
            mark("gather_keys")
            merged = self._merge_to_keys(allk, nq, k, st) if self.world > 1 else keys
            mark("merge")
            redo = torch.empty((nq,), dtype=torch.int32, device=dev)
            nredo = torch.zeros((1,), dtype=torch.int32, device=dev)
            L.check(lib.vdb_tq_check_dev(tq, C.c_void_p(merged.data_ptr()), k, n_total, C.c_void_p(tau.data_ptr()),
                                         C.c_void_p(ovf.data_ptr()), C.c_void_p(redo.data_ptr()),
                                         C.c_void_p(nredo.data_ptr())))
            mark("check")
        except Exception:
            lib.vdb_tq_end(tq)
            raise
        # every tensor the enqueued work touches stays referenced until _tensor_finish
        return {"tq": tq, "q": q, "merged": merged, "redo": redo, "nredo": nredo, "marks": marks,
                "keep": (jkeys, allj, tau, keys, ovf, allk)}

    def _tensor_finish(self, state, k):
        """After the chunk's stream has been joined: read the check result, redo flagged queries by the exact scan."""
        import torch
        lib = L.lib()
        try:
            merged, q = state["merged"], state["q"]
            st = C.c_void_p(torch.cuda.current_stream(q.device).cuda_stream)
            nr = int(state["nredo"].item())  # identical on every rank (same merged keys, same tau, reduced flags)
            marks = state["marks"]
            if marks:
                for (_, a), (name, b) in zip(marks[:-1], marks[1:]):
                    self.phase_ms[name] = self.phase_ms.get(name, 0.0) + a.elapsed_time(b)
                self.phase_ms["calls"] = self.phase_ms.get("calls", 0) + 1
            self.last_fallbacks += nr
            if nr:
                sel = torch.sort(state["redo"][:nr].long()).values
                rq = q.index_select(0, sel).contiguous()
                rk = self._scan_keys(rq, k, st)
                rall = self._gather(rk)
                rmerged = self._merge_to_keys(rall, nr, k, st) if self.world > 1 else rk
                merged.index_copy_(0, sel, rmerged)
            return merged
        finally:
            lib.vdb_tq_end(state["tq"])

    def _tensor_keys(self, q, k, info):
        """Global [nq, k] keys through the tensor-core phases (identical on every rank). Chunks of <= 16384 queries bound
        the scratch; the checks are read back once at the end. With VDB_GEMM_CHUNKS > 1 the chunks alternate between
        two side streams (tail of one chunk under the contraction of the next) - measured slower on 8 B200
        (6.23 / 6.55 / 7.05 ms per 10k-query step with 1 / 2 / 4 chunks), so it is off by default."""
        import torch
        nq, dev = q.shape[0], q.device
        self.last_fallbacks = 0
        per = -(-nq // max(1, self.PIPELINE_CHUNKS))             # ceil(nq / chunks)
        csize = nq if (nq < 4096 or self.PIPELINE_CHUNKS <= 1) else max(1024, -(-per // 256) * 256)  # whole 256-query tiles
        csize = min(csize, self.QUERY_CHUNK)
        bounds = [(c0, min(nq, c0 + csize)) for c0 in range(0, nq, csize)]
        if len(bounds) == 1:
            return self._tensor_finish(self._tensor_enqueue(q, k, info), k)
        cur = torch.cuda.current_stream(dev)
        if self._streams is None:
            self._streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
        states = []
        try:
            for s in self._streams:
                s.wait_stream(cur)
            for i, (c0, c1) in enumerate(bounds):
                with torch.cuda.stream(self._streams[i & 1]):
                    states.append(self._tensor_enqueue(q[c0:c1], k, info))
            for s in self._streams:
                cur.wait_stream(s)
            parts = []
            while states:
                parts.append(self._tensor_finish(states.pop(0), k))
            out = torch.cat(parts, 0)
            for t in parts:
                t.record_stream(cur)
            return out
        finally:
            for stt in states:  # only on an exception: release the remaining contexts once their work has drained
                torch.cuda.synchronize(dev)
                L.lib().vdb_tq_end(stt["tq"])

    # ---- search --------------------------------------------------------------------------------------------
    def knn_batch_dev(self, q, k):
        """q: torch CUDA tensor [nq, dim] (dataset dtype). Returns torch tensors (ids i64, dist f32, counts i32)."""
        import torch
        nq = q.shape[0]
        dev = q.device
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        lib = L.lib()
        ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
        dd = torch.empty((nq, k), dtype=torch.float32, device=dev)
        cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
        if self.world == 1:
            L.check(lib.vdb_flat_knn_dev(self.vec_set._h, C.c_void_p(q.data_ptr()), nq, k, C.c_void_p(ids.data_ptr()),
                                         C.c_void_p(dd.data_ptr()), C.c_void_p(cnt.data_ptr()), st))
            return ids, dd, cnt
        info = self._tensor_info(dev) if (nq >= self.TENSOR_MIN_NQ and 1 <= k <= 1024) else False
        if info:
            # chunks bound the candidate / rerank scratch of the filter phase
            merged = self._tensor_keys(q.contiguous(), k, info)
        else:
            merged = self._merge_to_keys(self._gather(self._scan_keys(q, k, st)), nq, k, st)
        L.check(lib.vdb_decode_keys_dev(C.c_void_p(merged.data_ptr()), nq, k, C.c_void_p(ids.data_ptr()),
                                        C.c_void_p(dd.data_ptr()), C.c_void_p(cnt.data_ptr()), st))
        return ids, dd, cnt

    def knn_batch(self, queries_pinned, k, out=None):
        """End-to-end call with HOST buffers: H2D of the queries, search, D2H of the results."""
        import torch
        dev = torch.device("cuda", torch.cuda.current_device())
        nq = queries_pinned.shape[0]
        if self.world > 1 and nq >= 64 * self.world:
            # every rank holds the same host batch: upload 1/world of it over this GPU's PCIe link and exchange the
            # slices over NVLink (one all-gather) instead of pushing the whole batch through every link
            import torch.distributed as dist
            per = -(-nq // self.world)
            lo, hi = min(nq, self.rank * per), min(nq, (self.rank + 1) * per)
            local = torch.zeros((per,) + tuple(queries_pinned.shape[1:]), dtype=queries_pinned.dtype, device=dev)
            local[:hi - lo].copy_(queries_pinned[lo:hi], non_blocking=True)
            q_all = torch.empty((self.world * per,) + tuple(queries_pinned.shape[1:]), dtype=queries_pinned.dtype, device=dev)
            dist.all_gather_into_tensor(q_all, local)
            q = q_all[:nq]
        else:
            q = queries_pinned.to(dev, non_blocking=True)
        ids, dd, cnt = self.knn_batch_dev(q, k)
        if out is None:
            out = (torch.empty(ids.shape, dtype=ids.dtype, pin_memory=True),
                   torch.empty(dd.shape, dtype=dd.dtype, pin_memory=True),
                   torch.empty(cnt.shape, dtype=cnt.dtype, pin_memory=True))
        out[0].copy_(ids, non_blocking=True)
        out[1].copy_(dd, non_blocking=True)
        out[2].copy_(cnt, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return out


class ShardedIVFIndex:
    """IVF over row shards: centroids replicated, every rank holds its slice of every list (members ascending),
    per-rank probe scans are merged by (distance, global id) after one all-gather."""

    def __init__(self, ivf_index, rank=0, world=1):
        self.ivf, self.rank, self.world = ivf_index, rank, world

    def knn_with_ef_batch_dev(self, q, k, n_probes):
        import torch
        nq, dev = q.shape[0], q.device
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        lib = L.lib()
        keys = torch.empty((nq, k), dtype=torch.int64, device=dev)
        L.check(lib.vdb_ivf_knn_keys_dev(self.ivf.vec_set._h, self.ivf._h, C.c_void_p(q.data_ptr()), nq, k, n_probes,
                                         C.c_void_p(keys.data_ptr()), st))
        return _gather_merge_decode(keys, nq, k, self.world, st)


class ShardedPQFlatIndex:
    """FlatIndex::knn_pq over row shards, identical to the unsharded call: the GLOBAL max(ef,k) best codes by ADC
    distance are selected first (all-gather + merge), then every rank reranks the candidates it owns."""

    def __init__(self, vec_set, pq_table, rank=0, world=1):
        self.vec_set, self.pq, self.rank, self.world = vec_set, pq_table, rank, world

    def knn_pq_batch_dev(self, q, k, ef):
        import torch
        nq, dev = q.shape[0], q.device
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        lib = L.lib()
        kk = max(ef, k)
        adc = torch.empty((nq, kk), dtype=torch.int64, device=dev)
        L.check(lib.vdb_pq_adc_keys_dev(self.vec_set._h, self.pq._h, C.c_void_p(q.data_ptr()), nq, kk,
                                        C.c_void_p(adc.data_ptr()), st))
        cand = _gather_merge_keys(adc, nq, kk, self.world, st)
        keys = torch.empty((nq, k), dtype=torch.int64, device=dev)
        L.check(lib.vdb_pq_rerank_keys_dev(self.vec_set._h, C.c_void_p(q.data_ptr()), nq, C.c_void_p(cand.data_ptr()), kk,
                                           k, C.c_void_p(keys.data_ptr()), st))
        return _gather_merge_decode(keys, nq, k, self.world, st)


class ReplicatedHNSWIndex:
    """HNSW does not shard (one walk depends on the previous hop): every rank holds the full
    vectors and graph, the query batch is split across the ranks and the results are all-gathered."""

    def __init__(self, hnsw_index, rank=0, world=1):
        self.hnsw, self.rank, self.world = hnsw_index, rank, world

    def knn_with_ef_batch_dev(self, q, k, ef):
        import torch
        nq, dev = q.shape[0], q.device
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        lib = L.lib()
        per = -(-nq // self.world)
        lo, hi = min(nq, self.rank * per), min(nq, (self.rank + 1) * per)
        ids = torch.full((per, k), -1, dtype=torch.int64, device=dev)
        dd = torch.full((per, k), float("nan"), dtype=torch.float32, device=dev)
        cnt = torch.zeros((per,), dtype=torch.int32, device=dev)
        if hi > lo:
            mine = q[lo:hi].contiguous()
            L.check(lib.vdb_hnsw_knn_dev(self.hnsw.vec_set._h, self.hnsw._h, C.c_void_p(mine.data_ptr()), hi - lo, k, ef,
                                         C.c_void_p(ids.data_ptr()), C.c_void_p(dd.data_ptr()), C.c_void_p(cnt.data_ptr()), st))
        return (_gather(ids, self.world).reshape(-1, k)[:nq], _gather(dd, self.world).reshape(-1, k)[:nq],
                _gather(cnt, self.world).reshape(-1)[:nq])


def _gather(t, world):
    import torch
    if world == 1:
        return t.unsqueeze(0)
    import torch.distributed as dist
    out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous())
    return out


def _gather_merge_keys(keys, nq, k, world, st):
    import torch
    if world == 1:
        return keys
    allk = _gather(keys, world)
    out = torch.empty((nq, k), dtype=torch.int64, device=keys.device)
    L.check(L.lib().vdb_merge_keys_to_keys_dev(C.c_void_p(allk.data_ptr()), world, nq, k, C.c_void_p(out.data_ptr()), st))
    return out


def _gather_merge_decode(keys, nq, k, world, st):
    import torch
    dev = keys.device
    allk = _gather(keys, world)
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    dd = torch.empty((nq, k), dtype=torch.float32, device=dev)
    cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
    L.check(L.lib().vdb_merge_keys_dev(C.c_void_p(allk.data_ptr()), world, nq, k, C.c_void_p(ids.data_ptr()),
                                       C.c_void_p(dd.data_ptr()), C.c_void_p(cnt.data_ptr()), st))
    return ids, dd, cnt
