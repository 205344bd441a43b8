"""Multi-GPU driver: one process per GPU (torch.distributed, NCCL over NVLink).  k-means runs on contiguous row
blocks; the search index is sharded either by ROWS (every GPU keeps the rows it already has = 1/G of every inverted
list) or by LISTS (default for G > 1: after k-means the rows are exchanged once with an all-to-all so every GPU owns
whole lists, balanced by size — the per-list work items keep their single-GPU size, so the scan scales with G
instead of shrinking every list to a few tiles).

  k-means : assign is embarrassingly parallel.  update needs Σ over ALL rows in row order (ivfflat.rs:52-55):
            reduce="chained" passes the running (sums, counts) from rank r-1 to rank r, which continues the
            left-to-right sum over its own rows, then the last rank broadcasts: the association is exactly the
            reference's, so centroids and assignments stay BIT-IDENTICAL to the CPU reference at any GPU count.
            reduce="allreduce" is the plain NCCL all-reduce of per-shard sums (fastest, association != reference's).
  search  : every GPU searches its shard for the whole query batch; the per-GPU top-k (ids+dists packed in one
            buffer) are all-gathered and merged by (distance, id) on every rank.
torch is used for what it is here for: device buffers, streams and the process group.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from ._abi import check, lib
from .index import Context, Dataset, IVFFlatIndex, KMeans


class _DevArray:
    """zero-copy torch view of library-owned device memory"""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(ptr, False), version=2)


def device_view(ptr: int, shape, dtype=torch.float32) -> torch.Tensor:
    typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.int64: "<i8", torch.uint8: "|u1"}[dtype]
    return torch.as_tensor(_DevArray(ptr, shape, typestr), device="cuda")


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_total: int, rank: int, world_size: int):
    per = (n_total + world_size - 1) // world_size
    r0 = min(n_total, rank * per)
    return r0, min(n_total, r0 + per) - r0


def chained_accumulate(tensors, local_step, group=None):
    """Ordered cross-rank accumulation: rank 0 zeroes `tensors`, every rank r > 0 first receives them from r-1, then
    `local_step()` continues the running values IN PLACE over its own rows, forwards them to r+1, and the last rank
    broadcasts the final values.  The association is that of one process walking all rows in order."""
    rank, ws = world()
    if rank == 0:
        for t in tensors:
            t.zero_()
    else:
        for t in tensors:
            dist.recv(t, src=rank - 1, group=group)
    local_step()
    if ws > 1:
        if rank < ws - 1:
            for t in tensors:
                dist.send(t, dst=rank + 1, group=group)
        for t in tensors:
            dist.broadcast(t, src=ws - 1, group=group)


def gather_init_centroids(rows: torch.Tensor, id_base: int, n_local: int, init_rows_global, ld: int, group=None):
    """initialize_centroids (ivfflat.rs:18-27) over row shards: the rank owning global row init[j] contributes it,
    everyone else zeros; an INTEGER all-reduce of the bit patterns is exact (and keeps -0.0, which a float add
    would turn into +0.0)."""
    _, ws = world()
    init = torch.as_tensor(np.ascontiguousarray(init_rows_global, np.int64), device=rows.device)
    local = init - id_base
    mine = (local >= 0) & (local < n_local)
    cents = torch.zeros((init.shape[0], ld), dtype=torch.float32, device=rows.device)
    if n_local:
        cents[mine] = rows[local[mine]]
    if ws > 1:
        ci = cents.view(torch.int32)
        dist.all_reduce(ci, op=dist.ReduceOp.SUM, group=group)
    return cents


def kmeans_fit_sharded(km: KMeans, init_rows_global: np.ndarray, max_iterations: int, reduce: str = "chained",
                       group=None) -> int:
    """IVFFlatIndex::build_kmeans (ivfflat.rs:73-100) over row shards.  init_rows_global are GLOBAL row numbers."""
    rank, ws = world()
    ds = km.ds
    Cn, ld = km.C, ds.ld
    dev = torch.device("cuda", torch.cuda.current_device())
    rows = device_view(ds.device_ptr, (max(ds.n, 1), ld))
    cents = gather_init_centroids(rows, ds.id_base, ds.n, init_rows_global, ld, group)

    def cur_centroids() -> torch.Tensor:
        p, l = C.c_void_p(), C.c_uint32()
        check(lib().vers_kmeans_centroids_device_ptr(km.h, C.byref(p), C.byref(l)))
        return device_view(p.value, (Cn, ld))

    cur_centroids().copy_(cents)
    sums = torch.zeros((Cn, ld), dtype=torch.float32, device=dev)
    counts = torch.zeros((Cn,), dtype=torch.int64, device=dev)
    it = 0
    while it < max_iterations:
        km.assign_step()
        if ws == 1 or reduce == "chained":
            chained_accumulate([sums, counts], lambda: km.sums_step_dev(sums.data_ptr(), counts.data_ptr()), group)
        elif reduce == "allreduce":
            sums.zero_(), counts.zero_()
            km.sums_step_dev(sums.data_ptr(), counts.data_ptr())
            dist.all_reduce(sums, group=group)
            dist.all_reduce(counts, group=group)
        else:
            raise ValueError(reduce)
        changed = km.finalize_step_dev(sums.data_ptr(), counts.data_ptr())
        it += 1
        if not changed:
            break
    km.assign_step()
    return it


def kmeans_cost_sharded(km: KMeans, group=None) -> np.float32:
    """calculate_kmeans_cost (ivfflat.rs:138-149) folded in global row order: rank r continues rank r-1's value"""
    rank, ws = world()
    dev = torch.device("cuda", torch.cuda.current_device())
    acc = torch.zeros(1, dtype=torch.float32, device=dev)
    if ws > 1 and rank > 0:
        dist.recv(acc, src=rank - 1, group=group)
    cost = km.cost_step(float(acc.item()))
    acc.fill_(float(cost))
    if ws > 1:
        if rank < ws - 1:
            dist.send(acc, dst=rank + 1, group=group)
        dist.broadcast(acc, src=ws - 1, group=group)
    return np.float32(acc.item())


def balanced_list_owners(sizes: np.ndarray, world_size: int) -> np.ndarray:
    """owner rank of every inverted list: largest list first onto the least-loaded rank (ties: lowest rank, lowest
    list), the same table on every rank"""
    order = np.lexsort((np.arange(sizes.shape[0]), -sizes.astype(np.int64)))
    load = np.zeros(world_size, np.int64)
    owner = np.zeros(sizes.shape[0], np.int64)
    for c in order:
        r = int(np.argmin(load))
        owner[c] = r
        load[r] += int(sizes[c])
    return owner


def exchange_rows_by_list(rows: torch.Tensor, assign: torch.Tensor, id_base: int, num_clusters: int, group=None):
    """The all-to-all of a list-sharded build, on whatever device the tensors live (NCCL on GPUs, gloo in the CPU
    tests).  rows [n, ld] fp32 and assign [n] int64 are this rank's contiguous block (global ids id_base ..).
    Returns (recv_rows, recv_ids int64, recv_assign int32, owner int64 [C]): the rows of the lists this rank owns, in
    ascending global id order (blocks arrive in rank order and ranks hold ascending id blocks)."""
    _, ws = world()
    dev = rows.device
    n, ld = rows.shape
    sizes = torch.bincount(assign, minlength=num_clusters)
    if ws > 1:
        dist.all_reduce(sizes, group=group)
    owner = torch.as_tensor(balanced_list_owners(sizes.cpu().numpy(), ws), device=dev)
    dest = owner[assign]
    order = torch.argsort(dest, stable=True)  # by destination, ascending local row (= ascending id) inside each
    send_counts = torch.bincount(dest, minlength=ws)
    recv_counts = torch.empty_like(send_counts)
    if ws > 1:
        dist.all_to_all_single(recv_counts, send_counts, group=group)
    else:
        recv_counts.copy_(send_counts)
    sc, rc = send_counts.cpu().tolist(), recv_counts.cpu().tolist()
    n_recv = int(sum(rc))

    def a2a(send: torch.Tensor) -> torch.Tensor:
        recv = torch.empty((n_recv,) + tuple(send.shape[1:]), dtype=send.dtype, device=dev)
        if ws > 1:
            dist.all_to_all_single(recv, send, rc, sc, group=group)
        else:
            recv.copy_(send)
        return recv

    recv_rows = a2a(rows.index_select(0, order))
    recv_ids = a2a(order + id_base)
    recv_assign = a2a(assign.index_select(0, order).to(torch.int32))
    return recv_rows, recv_ids, recv_assign, owner


def build_list_sharded(km: KMeans, group=None) -> IVFFlatIndex:
    """After k-means on row blocks: every rank sends each of its rows (with its global id and cluster) to the rank that
    owns the row's list (one all-to-all over NVLink) and builds its lists from what it receives.  Inside a list the
    rows stay in ascending id order like `ids[c]` (ivfflat.rs:123-127).  Lists a rank does not own are empty there;
    the centroid table is whole."""
    ds = km.ds
    Cn, ld, n = km.C, ds.ld, ds.n
    p = C.c_void_p()
    check(lib().vers_kmeans_assign_device_ptr(km.h, C.byref(p)))
    ds.ctx.sync()
    assign = device_view(p.value, (max(n, 1),), torch.int32)[:n].to(torch.int64)
    rows = device_view(ds.device_ptr, (max(n, 1), ld))[:n]
    recv_rows, recv_ids, recv_assign, _ = exchange_rows_by_list(rows, assign, ds.id_base, Cn, group)
    n_recv = recv_rows.shape[0]
    if n_recv == 0:  # keep the pointers valid
        recv_rows = torch.zeros((1, ld), dtype=torch.float32, device=rows.device)
        recv_ids = torch.zeros(1, dtype=torch.int64, device=rows.device)
        recv_assign = torch.zeros(1, dtype=torch.int32, device=rows.device)
    cp, cl = C.c_void_p(), C.c_uint32()
    check(lib().vers_kmeans_centroids_device_ptr(km.h, C.byref(cp), C.byref(cl)))
    torch.cuda.synchronize()
    local = Dataset.wrap_device(ds.ctx, recv_rows.data_ptr(), n_recv, ds.dim, 0, keepalive=recv_rows)
    ivf = IVFFlatIndex.from_parts_dev(local, cp.value, Cn, recv_assign.data_ptr(), recv_ids.data_ptr())
    ds.ctx.sync()
    # the list-major copy is built; the received row-major rows are no longer needed
    local.close()
    ivf._ds = None
    del recv_rows, local
    return ivf


class ShardedIVFFlat:
    """IVFFlatIndex whose rows are sharded over the ranks of the default process group."""

    def __init__(self, ivf: IVFFlatIndex, ctx: Context, peer_exchange: Optional[bool] = None):
        """peer_exchange: merge the per-GPU top-k with ONE kernel over NVLink peer memory (vers_peer_*: every rank
        stores its results into every peer's buffer, flags, waits, merges) instead of NCCL all-gather + merge kernel.
        Default: the environment variable VERS_PEER_GATHER=1 turns it on (opt-in this round: validated on 2 GPUs)."""
        self.ivf = ivf
        self.ctx = ctx
        self.rank, self.world = world()
        self._bufs = {}
        if peer_exchange is None:
            peer_exchange = os.environ.get("VERS_PEER_GATHER", "0") == "1"
        self._want_peer = bool(peer_exchange) and self.world > 1
        self._peer = None
        self._peer_slot = 0

    def _peer_handle(self, nq: int, k: int):
        """(re)creates the exchange buffers when the batch shape needs a larger slot; collective over all ranks"""
        need = nq * k * 12
        if self._peer is not None and need <= self._peer_slot:
            return self._peer
        if self._peer is not None:
            torch.cuda.synchronize()
            dist.barrier()
            check(lib().vers_peer_free(self._peer))
            self._peer = None
        h = C.c_void_p()
        mine = (C.c_ubyte * 64)()
        check(lib().vers_peer_create(self.ctx.h, self.world, self.rank, need, C.byref(h), mine))
        dev = torch.device("cuda", torch.cuda.current_device())
        t = torch.tensor(list(mine), dtype=torch.uint8, device=dev)
        allh = torch.empty(self.world * 64, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh, t)
        raw = bytes(allh.cpu().numpy().tobytes())
        buf = (C.c_ubyte * len(raw)).from_buffer_copy(raw)
        check(lib().vers_peer_connect(h, buf))
        torch.cuda.synchronize()
        dist.barrier()  # every rank has mapped every buffer before anyone stores into a peer
        self._peer, self._peer_slot = h, need
        return h

    @classmethod
    def build(cls, ds: Dataset, num_clusters: int, max_iterations: int, init_rows_global: np.ndarray,
              reduce: str = "chained", shard_by: str = "lists") -> "ShardedIVFFlat":
        km = KMeans(ds, num_clusters)
        kmeans_fit_sharded(km, init_rows_global, max_iterations, reduce)
        _, ws = world()
        if ws == 1 or shard_by == "rows":
            ivf = IVFFlatIndex.from_kmeans(km)
        elif shard_by == "lists":
            ivf = build_list_sharded(km)
        else:
            raise ValueError(shard_by)
        km.close()
        return cls(ivf, ds.ctx)

    def _buffers(self, nq: int, k: int):
        key = (nq, k)
        if key not in self._bufs:
            dev = torch.device("cuda", torch.cuda.current_device())
            nk = nq * k
            L = nk + (nk + 1) // 2  # int64 words: nk ids, then nk floats packed two per word
            local = torch.empty(L, dtype=torch.int64, device=dev)
            allb = torch.empty((self.world, L), dtype=torch.int64, device=dev)
            out_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
            out_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
            out_c = torch.empty((nq,), dtype=torch.int32, device=dev)
            self._bufs[key] = (local, allb, out_ids, out_d, out_c, L)
        return self._bufs[key]

    def capture_search(self, d_queries: torch.Tensor, top_k: int, nprobe: int):
        """Single GPU: captures one search_dev call on this batch buffer into a CUDA graph (the ~45 small launches of
        a step become one graph launch).  Returns (graph, outputs): refill ``d_queries`` in place and
        ``graph.replay()``; the outputs are the same device tensors every time.  The library makes no host
        synchronisation and no allocation on this path once it has run eagerly with the same shapes."""
        if self.world > 1:
            raise RuntimeError("capture_search: capturing the NCCL all-gathers of the multi-GPU step is not supported "
                               "(it deadlocked with torch 2.11 / NCCL 2.28 in this image)")
        for _ in range(2):
            self.search_dev(d_queries, top_k, nprobe)
        torch.cuda.synchronize()
        eager_stream = torch.cuda.current_stream()
        g = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(g):
                self.ctx.set_stream(torch.cuda.current_stream().cuda_stream)
                out = self.search_dev(d_queries, top_k, nprobe)
        finally:
            self.ctx.set_stream(eager_stream.cuda_stream)
        return g, out

    def search_dev(self, d_queries: torch.Tensor, top_k: int, nprobe: int):
        """d_queries: [nq, ld] float32 on this rank's GPU (the same batch on every rank).  Returns device tensors
        (ids int64 [nq,k] holding u64 bit patterns, dists [nq,k], counts [nq]) — the global result on every rank."""
        nq = d_queries.shape[0]
        local, allb, out_ids, out_d, out_c, L = self._buffers(nq, top_k)
        nk = nq * top_k
        ids_ptr = local.data_ptr()
        d_ptr = ids_ptr + nk * 8
        if self.world == 1:
            self.ivf.search_batch_dev(d_queries.data_ptr(), nq, top_k, nprobe, out_ids.data_ptr(), out_d.data_ptr(),
                                      out_c.data_ptr())
            return out_ids, out_d, out_c
        # the centroid probe is replicated work: every rank probes 1/world of the batch, the probe lists are
        # all-gathered (nq x nprobe x 8 B), then every rank scans its row shard of exactly those lists
        per = (nq + self.world - 1) // self.world
        npb = min(nprobe, self.ivf.num_centroids)
        key = ("probe", nq, npb)
        if key not in self._bufs:
            dev = d_queries.device
            self._bufs[key] = (torch.full((per, npb), -1, dtype=torch.int64, device=dev),
                               torch.empty((self.world * per, npb), dtype=torch.int64, device=dev))
        p_local, p_all = self._bufs[key]
        q0 = min(nq, self.rank * per)
        nql = max(0, min(nq, q0 + per) - q0)
        if nql:
            check(lib().vers_ivf_probe_dev(self.ivf.h, C.c_void_p(d_queries.data_ptr() + q0 * d_queries.shape[1] * 4), nql,
                                           npb, C.c_void_p(p_local.data_ptr())))
        dist.all_gather_into_tensor(p_all, p_local)
        check(lib().vers_ivf_search_probed_dev(self.ivf.h, C.c_void_p(d_queries.data_ptr()), nq, top_k, npb,
                                               C.c_void_p(p_all.data_ptr()), C.c_void_p(ids_ptr), C.c_void_p(d_ptr),
                                               C.c_void_p(out_c.data_ptr())))
        if self._want_peer:
            # exchange + merge as one kernel over NVLink peer memory
            peer = self._peer_handle(nq, top_k)
            check(lib().vers_peer_gather_merge_dev(peer, C.c_void_p(ids_ptr), C.c_void_p(d_ptr), nq, top_k,
                                                   C.c_void_p(out_ids.data_ptr()), C.c_void_p(out_d.data_ptr()),
                                                   C.c_void_p(out_c.data_ptr())))
            return out_ids, out_d, out_c
        dist.all_gather_into_tensor(allb, local)
        base = allb.data_ptr()
        check(lib().vers_topk_merge_dev(self.ctx.h, C.c_void_p(base), C.c_void_p(base + nk * 8), self.world, L, 2 * L,
                                        nq, top_k, C.c_void_p(out_ids.data_ptr()), C.c_void_p(out_d.data_ptr()),
                                        C.c_void_p(out_c.data_ptr())))
        return out_ids, out_d, out_c
