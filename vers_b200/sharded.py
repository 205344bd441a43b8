"""Multi-GPU driver: one process per GPU.  Everything that touches the data path lives behind the C ABI
(`vers_comm_*`, `vers_sharded_*` in include/vers_device.h, csrc/comm.cu): NCCL bootstrap, k-means over row shards
(chained ordered reduction or all-reduce), the all-to-all that shards the index by inverted list, and the search step
whose two exchanges run over NVLink peer memory.  This module only (a) ships NCCL's 128-byte unique id between the
processes (torch.distributed's store when a process group exists — any transport would do) and (b) offers the small
torch conveniences bench.py and the tests use (zero-copy views of library memory, CUDA-graph capture).

  k-means : assign is embarrassingly parallel.  update needs Σ over ALL rows in row order (ivfflat.rs:52-55):
            reduce="chained" passes the running (sums, counts) from rank r-1 to rank r, which continues the
            left-to-right sum over its own rows, then the last rank broadcasts: the association is exactly the
            reference's, so centroids and assignments stay BIT-IDENTICAL to the CPU reference at any GPU count.
            reduce="allreduce" is the plain NCCL all-reduce of per-shard sums (fastest, association != reference's).
  search  : every rank probes 1/world of the batch, the probe lists are all-gathered, every rank scans the lists it
            owns, the per-GPU top-k are exchanged and merged by (distance, id) — vers_sharded_ivf_search[_dev].
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from ._abi import REDUCE_ALLREDUCE, REDUCE_CHAINED, check, lib
from .index import Context, Dataset, IVFFlatIndex, KMeans


def world():
    """(rank, world size) of this process: torch.distributed's if initialised, else torchrun's environment"""
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_bounds(n_total: int, rank: int, world_size: int):
    per = (n_total + world_size - 1) // world_size
    r0 = min(n_total, rank * per)
    return r0, min(n_total, r0 + per) - r0


class _DevArray:
    """zero-copy torch view of library-owned device memory"""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(ptr, False), version=2)


def device_view(ptr: int, shape, dtype=None):
    import torch

    dtype = dtype or torch.float32
    typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.int64: "<i8", torch.uint8: "|u1"}[dtype]
    return torch.as_tensor(_DevArray(ptr, shape, typestr), device="cuda")


class Comm:
    """vers_comm: this rank's membership in the group of GPUs (include/vers_device.h).  The 128-byte NCCL unique id
    travels through torch.distributed (broadcast_object_list on whatever backend the group has: gloo or nccl)."""

    def __init__(self, ctx: Context, rank: Optional[int] = None, world_size: Optional[int] = None,
                 unique_id: Optional[bytes] = None):
        r, w = world()
        self.rank = r if rank is None else rank
        self.world = w if world_size is None else world_size
        self.ctx = ctx
        idbuf = None
        if self.world > 1:
            if unique_id is None:
                import torch.distributed as dist

                box = [None]
                if self.rank == 0:
                    mine = (C.c_ubyte * 128)()
                    check(lib().vers_comm_unique_id(mine))
                    box[0] = bytes(mine)
                dist.broadcast_object_list(box, src=0)
                unique_id = box[0]
            idbuf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        h = C.c_void_p()
        check(lib().vers_comm_create(ctx.h, self.world, self.rank, idbuf, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            check(lib().vers_comm_destroy(self.h))
            self.h = None

    def barrier(self):
        check(lib().vers_comm_barrier(self.h))

    def max_over_ranks(self, x: float) -> float:
        v = C.c_double(x)
        check(lib().vers_comm_max_f64(self.h, C.byref(v)))
        return float(v.value)

    @property
    def last_exchange_s(self) -> float:
        v = C.c_double()
        check(lib().vers_comm_info(self.h, None, None, C.byref(v)))
        return float(v.value)


_REDUCE = {"chained": REDUCE_CHAINED, "allreduce": REDUCE_ALLREDUCE}


def kmeans_fit_sharded(comm: Comm, km: KMeans, init_rows_global: np.ndarray, max_iterations: int,
                       reduce: str = "chained") -> int:
    """IVFFlatIndex::build_kmeans (ivfflat.rs:73-100) over row shards.  init_rows_global are GLOBAL row numbers."""
    init = np.ascontiguousarray(init_rows_global, np.uint64)
    if init.shape != (km.C,):
        raise ValueError("init_rows_global must hold num_clusters row numbers")
    it = C.c_uint32()
    check(lib().vers_sharded_kmeans_fit(comm.h, km.h, init.ctypes.data_as(C.c_void_p), max_iterations, _REDUCE[reduce],
                                        C.byref(it)))
    return int(it.value)


def kmeans_cost_sharded(comm: Comm, km: KMeans) -> np.float32:
    """calculate_kmeans_cost (ivfflat.rs:138-149) folded in global row order: rank r continues rank r-1's value"""
    c = C.c_float()
    check(lib().vers_sharded_kmeans_cost(comm.h, km.h, C.byref(c)))
    return np.float32(c.value)


def balanced_list_owners(sizes: np.ndarray, world_size: int) -> np.ndarray:
    """owner rank of every inverted list: largest list first onto the least-loaded rank (ties: lowest rank, lowest
    list) — the table vers_sharded_ivf_build uses (vers_sharded_list_owners: host arithmetic, no GPU needed)"""
    sz = np.ascontiguousarray(sizes, np.uint64)
    owner = np.empty(sz.shape[0], np.uint32)
    check(lib().vers_sharded_list_owners(sz.ctypes.data_as(C.c_void_p), sz.shape[0], world_size,
                                         owner.ctypes.data_as(C.c_void_p)))
    return owner.astype(np.int64)


def build_list_sharded(comm: Comm, km: KMeans) -> IVFFlatIndex:
    """After k-means on row blocks: every rank sends each of its rows (with its global id and cluster) to the rank that
    owns the row's list (one all-to-all over NVLink) and builds its lists from what it receives."""
    h = C.c_void_p()
    check(lib().vers_sharded_ivf_build(comm.h, km.h, C.byref(h)))
    return IVFFlatIndex(km.ds.ctx, h, None, None)


class ShardedIVFFlat:
    """IVFFlatIndex sharded over the ranks of a Comm (by inverted list, or by rows when built with shard_by="rows")."""

    def __init__(self, ivf: IVFFlatIndex, comm: Comm):
        self.ivf = ivf
        self.comm = comm
        self.ctx = comm.ctx
        self.rank, self.world = comm.rank, comm.world
        self._bufs = {}

    @classmethod
    def build(cls, comm: Comm, ds: Dataset, num_clusters: int, max_iterations: int, init_rows_global: np.ndarray,
              reduce: str = "chained", shard_by: str = "lists") -> "ShardedIVFFlat":
        km = KMeans(ds, num_clusters)
        kmeans_fit_sharded(comm, km, init_rows_global, max_iterations, reduce)
        if comm.world == 1 or shard_by == "rows":
            ivf = IVFFlatIndex.from_kmeans(km)
        elif shard_by == "lists":
            ivf = build_list_sharded(comm, km)
        else:
            raise ValueError(shard_by)
        km.close()
        return cls(ivf, comm)

    def search(self, queries: np.ndarray, top_k: int, nprobe: int):
        """host buffers in and out through vers_sharded_ivf_search (every rank passes the same batch)"""
        q = np.ascontiguousarray(queries, np.float32)
        nq = q.shape[0]
        ids = np.empty((nq, top_k), np.uint64)
        d = np.empty((nq, top_k), np.float32)
        cnt = np.empty(nq, np.uint32)
        check(lib().vers_sharded_ivf_search(self.comm.h, self.ivf.h, q.ctypes.data_as(C.c_void_p), nq, q.shape[1], top_k,
                                            nprobe, ids.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p),
                                            cnt.ctypes.data_as(C.c_void_p)))
        return ids, d, cnt

    def _buffers(self, nq: int, k: int, device):
        import torch

        key = (nq, k)
        if key not in self._bufs:
            self._bufs[key] = (torch.empty((nq, k), dtype=torch.int64, device=device),
                               torch.empty((nq, k), dtype=torch.float32, device=device),
                               torch.empty((nq,), dtype=torch.int32, device=device))
        return self._bufs[key]

    def search_dev(self, d_queries, top_k: int, nprobe: int):
        """d_queries: [nq, ld] float32 torch tensor on this rank's GPU (the same batch on every rank).  Returns device
        tensors (ids int64 [nq,k] holding u64 bit patterns, dists [nq,k], counts [nq]) — the global result on every
        rank.  Enqueued on the context's stream, no host synchronisation."""
        nq = d_queries.shape[0]
        out_ids, out_d, out_c = self._buffers(nq, top_k, d_queries.device)
        check(lib().vers_sharded_ivf_search_dev(self.comm.h, self.ivf.h, C.c_void_p(d_queries.data_ptr()), nq, top_k,
                                                nprobe, C.c_void_p(out_ids.data_ptr()), C.c_void_p(out_d.data_ptr()),
                                                C.c_void_p(out_c.data_ptr())))
        return out_ids, out_d, out_c

    def capture_search(self, d_queries, top_k: int, nprobe: int):
        """Captures one search_dev call on this batch buffer into a CUDA graph (the ~45 small launches of a step become
        one graph launch).  Works at any GPU count: the step has no NCCL call in it, and the exchange kernels keep their
        step counters in device memory, so a replay advances the peer protocol like an eager step.  Every rank must
        capture and replay in lockstep.  Returns (graph, outputs): refill ``d_queries`` in place and ``graph.replay()``."""
        import torch

        for _ in range(2):
            self.search_dev(d_queries, top_k, nprobe)
        torch.cuda.synchronize()
        self.comm.barrier()
        eager_stream = torch.cuda.current_stream()
        g = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(g):
                self.ctx.set_stream(torch.cuda.current_stream().cuda_stream)
                out = self.search_dev(d_queries, top_k, nprobe)
        finally:
            self.ctx.set_stream(eager_stream.cuda_stream)
        return g, out
