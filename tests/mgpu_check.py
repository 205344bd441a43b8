"""Multi-GPU parity check, launched under torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
Every rank holds a contiguous block of the rows.  Everything goes through the C ABI (vers_comm_* / vers_sharded_*,
csrc/comm.cu).  With reduce = chained the sharded k-means must reproduce the SINGLE-PROCESS oracle bit for bit
(assignments, centroid bits, cost bits); the sharded search (row shards and list shards; device buffers, host buffers,
CUDA-graph replays, changing batch shapes) must return the oracle's ids and distance bits.  reduce = allreduce is
checked against the oracle's sharded-order mode only up to tolerance (NCCL's reduction order is its own)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as vo  # noqa: E402
import vers_b200 as vb  # noqa: E402
from vers_b200.sharded import (Comm, ShardedIVFFlat, build_list_sharded, kmeans_cost_sharded, kmeans_fit_sharded,  # noqa: E402
                               shard_bounds)

bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)


def main():
    rank, ws, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("gloo")  # only carries the NCCL unique id: the data path is the library's own
    ctx = vb.Context(lr)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    comm = Comm(ctx)
    n, dim, C, k = 20011, 300, 48, 10
    rows = vo.synth(1, n, dim, kind=1, n_centers=40, center_seed=7)
    q = vo.synth(2, 64, dim, kind=1, n_centers=40, center_seed=7)
    init = vo.init_rows(3, 1, C, n)[0]
    r0, nl = shard_bounds(n, rank, ws)
    ds = vb.Dataset.synth(ctx, 1, nl, dim, kind=1, n_centers=40, center_seed=7, row0=r0)
    assert np.array_equal(bits(ds.download()), bits(rows[r0:r0 + nl]))

    # --- chained k-means == the single-process reference order
    km = vb.KMeans(ds, C)
    iters = kmeans_fit_sharded(comm, km, init, 12, reduce="chained")
    cents, assign, oit = vo.kmeans_fit(rows, init, 12)
    assert iters == oit, (iters, oit)
    assert np.array_equal(km.assignments(), assign[r0:r0 + nl]), "sharded assignments differ from the oracle"
    assert np.array_equal(bits(km.centroids()), bits(cents)), "sharded centroid bits differ from the oracle"
    cost = kmeans_cost_sharded(comm, km)
    assert bits(np.float32(cost)) == bits(vo.kmeans_cost(rows, cents, assign)), "chained cost differs"

    # --- sharded search (rows: every GPU holds 1/G of every list; lists: every GPU owns whole lists after an
    #     all-to-all of the rows): peer-memory exchange + merge == oracle
    off, lrw = vo.ivf_lists(assign, C)
    d_q = torch.from_numpy(np.ascontiguousarray(q)).cuda()
    for shard_by in ("rows", "lists"):
        ivf = vb.IVFFlatIndex.from_kmeans(km) if shard_by == "rows" else build_list_sharded(comm, km)
        index = ShardedIVFFlat(ivf, comm)
        if shard_by == "lists":
            sizes = ivf.list_sizes.astype(np.int64)
            want_sizes = np.bincount(assign.astype(np.int64), minlength=C)
            owner = vb.sharded.balanced_list_owners(want_sizes, ws)
            assert np.array_equal(sizes, np.where(owner == rank, want_sizes, 0)), "list ownership / sizes differ"
            if ws <= C:
                assert 0 < int((sizes > 0).sum()) < C or ws == 1, "every rank must own some lists, not all"
            c0 = int(np.flatnonzero(owner == rank)[0])  # ids of an owned list: ascending global ids (ivfflat.rs:123-127)
            lids = ivf.get_list(c0)
            assert np.array_equal(lids, np.flatnonzero(assign == c0).astype(np.uint64))
        for mode, nprobe in [(m, p) for m in (0, 4) for p in (1, 8, C)]:
            # mode 0: candidates from the fp32 rows (split tf32); mode 4: from the fp16 candidate copy of this rank's lists
            ivf.set_mode(mode)
            oi, od, oc = vo.ivf_search(rows, cents, off, lrw, q, k, nprobe=nprobe)
            for rep in range(3):  # repeated steps: the double-buffered slots and flags of the peer protocol
                ids, d, cnt = index.search_dev(d_q, k, nprobe)
            torch.cuda.synchronize()
            assert np.array_equal(ids.cpu().numpy().view(np.uint64), oi), f"ids differ nprobe={nprobe} {shard_by} {mode}"
            assert np.array_equal(bits(d.cpu().numpy()), bits(od)), f"distances differ nprobe={nprobe} {shard_by} {mode}"
            assert np.array_equal(cnt.cpu().numpy().astype(np.uint32), oc)
            # host buffers through vers_sharded_ivf_search, an odd batch size (ranks get unequal probe shares)
            hi, hd, hc = index.search(q[:37], k, nprobe)
            assert np.array_equal(hi, oi[:37]) and np.array_equal(bits(hd), bits(od[:37])) and np.array_equal(hc, oc[:37])
        # a larger top_k / batch than the exchange buffers were sized for: they are re-created collectively
        big_q = np.vstack([q, q[::-1]])
        oi, od, oc = vo.ivf_search(rows, cents, off, lrw, big_q, 40, nprobe=8)
        hi, hd, hc = index.search(big_q, 40, 8)
        assert np.array_equal(hi, oi) and np.array_equal(bits(hd), bits(od)) and np.array_equal(hc, oc)
        # CUDA-graph replays of the whole step (no NCCL inside; the step counters live in device memory)
        g, out = index.capture_search(d_q, k, 8)
        oi, od, oc = vo.ivf_search(rows, cents, off, lrw, q, k, nprobe=8)
        for rep in range(4):
            d_q.copy_(torch.from_numpy(np.ascontiguousarray(q[::-1] if rep % 2 else q)).cuda())
            g.replay()
        torch.cuda.synchronize()
        assert np.array_equal(out[0].cpu().numpy().view(np.uint64), oi[::-1]), "graph replay ids differ"
        assert np.array_equal(bits(out[1].cpu().numpy()), bits(od[::-1]))
        d_q.copy_(torch.from_numpy(np.ascontiguousarray(q)).cuda())
        ids, d, cnt = index.search_dev(d_q, k, 8)  # and eager steps after replays keep the protocol in step
        torch.cuda.synchronize()
        assert np.array_equal(ids.cpu().numpy().view(np.uint64), oi)
        del g

    # --- hyperplane forest: replicated on every rank (tree construction does not shard), queries sharded, slices
    #     all-gathered (vers_sharded_lsh_search): the oracle's ids and distance bits on every rank
    import ctypes
    frows = rows[:6000]
    g = vb.ANNIndex.build_index(6, 40, frows, None, seed=4, ctx=ctx)
    o = vo.LSH(frows, None, 6, 40, 4)
    fq = np.ascontiguousarray(q[:51])
    fi = np.empty((51, 7), np.uint64)
    fd = np.empty((51, 7), np.float32)
    fc = np.empty(51, np.uint32)
    vp = ctypes.c_void_p
    vb._abi.check(vb.lib().vers_sharded_lsh_search(comm.h, g.h, fq.ctypes.data_as(vp), 51, dim, 7, fi.ctypes.data_as(vp),
                                                   fd.ctypes.data_as(vp), fc.ctypes.data_as(vp)))
    oi, od, oc = o.search(fq, 7)
    assert np.array_equal(fi, oi) and np.array_equal(bits(fd), bits(od)) and np.array_equal(fc, oc), "sharded LSH search"

    # --- all-reduce mode: same counts, centroids within tolerance of the oracle's sharded-order mode
    km2 = vb.KMeans(ds, C)
    kmeans_fit_sharded(comm, km2, init, 1, reduce="allreduce")
    c0 = rows[init.astype(np.int64)]
    a0 = vo.assign(rows, c0)
    want, _ = vo.update(rows, a0, C, shards=ws)
    assert np.allclose(km2.centroids(), want, rtol=1e-5, atol=1e-7)
    comm.barrier()
    if rank == 0:
        print(f"mgpu_check ok: world={ws}, chained k-means bit-identical to the single-process oracle "
              f"({iters} iterations), sharded search (row shards and list shards; fp32-row and fp16-copy candidate "
              f"passes; device, host and graph-replay paths) ids+distances identical, sharded LSH search identical")
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
