"""Multi-GPU parity check, launched under torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
Every rank holds a contiguous block of the rows.  With reduce="chained" the sharded k-means must reproduce the
SINGLE-PROCESS oracle bit for bit (assignments, centroid bits), and the all-gather + merge search must return the
oracle's ids and distance bits.  reduce="allreduce" is checked against the oracle's sharded-order mode only up to
tolerance (NCCL's reduction order is its own)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as vo  # noqa: E402
import vers_b200 as vb  # noqa: E402
from vers_b200.sharded import (ShardedIVFFlat, build_list_sharded, kmeans_cost_sharded, kmeans_fit_sharded,  # noqa: E402
                               shard_bounds)

bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)


def main():
    rank, ws, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ctx = vb.Context(lr)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    n, dim, C, k = 20011, 300, 48, 10
    rows = vo.synth(1, n, dim, kind=1, n_centers=40, center_seed=7)
    q = vo.synth(2, 64, dim, kind=1, n_centers=40, center_seed=7)
    init = vo.init_rows(3, 1, C, n)[0]
    r0, nl = shard_bounds(n, rank, ws)
    ds = vb.Dataset.synth(ctx, 1, nl, dim, kind=1, n_centers=40, center_seed=7, row0=r0)
    assert np.array_equal(bits(ds.download()), bits(rows[r0:r0 + nl]))

    # --- chained k-means == the single-process reference order
    km = vb.KMeans(ds, C)
    iters = kmeans_fit_sharded(km, init, 12, reduce="chained")
    cents, assign, oit = vo.kmeans_fit(rows, init, 12)
    assert iters == oit, (iters, oit)
    assert np.array_equal(km.assignments(), assign[r0:r0 + nl]), "sharded assignments differ from the oracle"
    assert np.array_equal(bits(km.centroids()), bits(cents)), "sharded centroid bits differ from the oracle"
    cost = kmeans_cost_sharded(km)
    assert bits(np.float32(cost)) == bits(vo.kmeans_cost(rows, cents, assign)), "chained cost differs"

    # --- sharded search (rows: every GPU holds 1/G of every list; lists: every GPU owns whole lists after an
    #     all-to-all of the rows): all-gather + merge == oracle
    off, lrw = vo.ivf_lists(assign, C)
    d_q = torch.from_numpy(np.ascontiguousarray(q)).cuda()
    for shard_by in ("rows", "lists"):
        ivf = vb.IVFFlatIndex.from_kmeans(km) if shard_by == "rows" else build_list_sharded(km)
        index = ShardedIVFFlat(ivf, ctx, peer_exchange=(shard_by == "lists" and os.environ.get("VERS_PEER_GATHER") == "1"))
        if shard_by == "lists":
            sizes = torch.as_tensor(ivf.list_sizes.astype(np.int64)).cuda()
            dist.all_reduce(sizes)
            assert np.array_equal(sizes.cpu().numpy(), np.bincount(assign.astype(np.int64), minlength=C))
            owned = int((ivf.list_sizes > 0).sum())
            assert 0 < owned < C, "every rank must own some lists, not all"
        for nprobe in (1, 8, C):
            ids, d, cnt = index.search_dev(d_q, k, nprobe)
            torch.cuda.synchronize()
            oi, od, oc = vo.ivf_search(rows, cents, off, lrw, q, k, nprobe=nprobe)
            assert np.array_equal(ids.cpu().numpy().view(np.uint64), oi), f"ids differ nprobe={nprobe} {shard_by}"
            assert np.array_equal(bits(d.cpu().numpy()), bits(od)), f"distances differ nprobe={nprobe} {shard_by}"
            assert np.array_equal(cnt.cpu().numpy().astype(np.uint32), oc)

    # --- all-reduce mode: same counts, centroids within tolerance of the oracle's sharded-order mode
    km2 = vb.KMeans(ds, C)
    kmeans_fit_sharded(km2, init, 1, reduce="allreduce")
    c0 = rows[init.astype(np.int64)]
    a0 = vo.assign(rows, c0)
    want, _ = vo.update(rows, a0, C, shards=ws)
    assert np.allclose(km2.centroids(), want, rtol=1e-5, atol=1e-7)
    dist.barrier()
    if rank == 0:
        print(f"mgpu_check ok: world={ws}, chained k-means bit-identical to the single-process oracle "
              f"({iters} iterations), sharded search (row shards and list shards) ids+distances identical")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
