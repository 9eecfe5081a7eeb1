"""Row-sharded Flat search: one process per GPU, database rows split into contiguous blocks,
per-GPU top-k merged after one NCCL all-gather of [nq, k] packed keys.

torch is used here only as plumbing (device buffers, streams, torch.distributed); every compute
step is a kernel of libvdb_b200.so.
"""
import ctypes as C

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
    """Flat index over this rank's row block; search results are global and identical on all ranks."""

    def __init__(self, vec_set, rank=0, world=1):
        self.vec_set = vec_set
        self.rank, self.world = rank, world

    def knn_batch_dev(self, q, k):
        """q: torch CUDA tensor [nq, dim] (dataset dtype). Returns torch tensors (ids i64, dist f32, counts i32)."""
        import torch
        nq = q.shape[0]
        dev = q.device
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        keys = torch.empty((nq, k), dtype=torch.int64, device=dev)
        lib = L.lib()
        L.check(lib.vdb_flat_knn_keys_dev(self.vec_set._h, C.c_void_p(q.data_ptr()), nq, k,
                                          C.c_void_p(keys.data_ptr()), st))
        if self.world > 1:
            import torch.distributed as dist
            allk = torch.empty((self.world, nq, k), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(allk, keys)
            nlists = self.world
        else:
            allk, nlists = keys, 1
        ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
        dd = torch.empty((nq, k), dtype=torch.float32, device=dev)
        cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
        L.check(lib.vdb_merge_keys_dev(C.c_void_p(allk.data_ptr()), nlists, nq, k, C.c_void_p(ids.data_ptr()),
                                       C.c_void_p(dd.data_ptr()), C.c_void_p(cnt.data_ptr()), st))
        return ids, dd, cnt

    def knn_batch(self, queries_pinned, k, out=None):
        """End-to-end call with HOST buffers: H2D of the queries, search, D2H of the results."""
        import torch
        dev = torch.device("cuda", torch.cuda.current_device())
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
