"""CPU model (torch + gloo) of the multi-rank protocols that vers_b200/csrc/comm.cu runs over NCCL on the GPUs — test
infrastructure, like the CPU model of the warp selection networks.  The per-shard arithmetic is supplied by the caller
(the oracle in the tests); what is modelled is WHO sends WHAT to WHOM in WHICH order:

  chained_accumulate      : vers_sharded_kmeans_fit, reduce = VERS_REDUCE_CHAINED (ncclRecv from rank-1, local sums,
                            ncclSend to rank+1, ncclBroadcast from the last rank)
  gather_init_centroids   : its centroid initialisation (owner contributes the row, INTEGER all-reduce of the bits)
  exchange_rows_by_list   : vers_sharded_ivf_build (owner table, stable order by destination, counts matrix,
                            all-to-all of rows / global ids / clusters)
"""
import numpy as np
import torch
import torch.distributed as dist


def _world():
    return dist.get_rank(), dist.get_world_size()


def chained_accumulate(tensors, local_step):
    rank, ws = _world()
    if rank == 0:
        for t in tensors:
            t.zero_()
    else:
        for t in tensors:
            dist.recv(t, src=rank - 1)
    local_step()
    if ws > 1:
        if rank < ws - 1:
            for t in tensors:
                dist.send(t, dst=rank + 1)
        for t in tensors:
            dist.broadcast(t, src=ws - 1)


def gather_init_centroids(rows: torch.Tensor, id_base: int, n_local: int, init_rows_global, ld: int):
    init = torch.as_tensor(np.ascontiguousarray(init_rows_global, np.int64))
    local = init - id_base
    mine = (local >= 0) & (local < n_local)
    cents = torch.zeros((init.shape[0], ld), dtype=torch.float32)
    if n_local:
        cents[mine] = rows[local[mine]]
    dist.all_reduce(cents.view(torch.int32), op=dist.ReduceOp.SUM)
    return cents


def exchange_rows_by_list(rows: torch.Tensor, assign: torch.Tensor, id_base: int, num_clusters: int, owners_fn):
    """rows [n, ld] fp32 and assign [n] int64: this rank's contiguous block (global ids id_base ..).  Returns
    (recv_rows, recv_ids int64, recv_assign int64, owner int64 [C])."""
    _, ws = _world()
    sizes = torch.bincount(assign, minlength=num_clusters)
    dist.all_reduce(sizes)
    owner = torch.as_tensor(owners_fn(sizes.numpy(), ws))
    dest = owner[assign]
    order = torch.argsort(dest, stable=True)  # by destination, ascending local row (= ascending id) inside each
    send_counts = torch.bincount(dest, minlength=ws)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts)
    sc, rc = send_counts.tolist(), recv_counts.tolist()

    def a2a(send: torch.Tensor) -> torch.Tensor:
        recv = torch.empty((int(sum(rc)),) + tuple(send.shape[1:]), dtype=send.dtype)
        dist.all_to_all_single(recv, send, rc, sc)
        return recv

    return a2a(rows.index_select(0, order)), a2a(order + id_base), a2a(assign.index_select(0, order)), owner
