"""N>1 logic on CPU: world_size-2 gloo process group (no GPU).  The multi-GPU data path lives in csrc/comm.cu (NCCL +
peer memory); what runs here is (a) a torch/gloo MODEL of its protocols (tests/gloo_model.py: rank-chained ordered
accumulation that keeps multi-GPU k-means bit-identical to the single-process reference order, the exact
initial-centroid exchange, the all-to-all of a list-sharded build) with the oracle supplying the per-shard
arithmetic, (b) the host arithmetic the library exposes without a GPU (vers_sharded_list_owners) against that model,
(c) the bootstrap of vers_b200.sharded.Comm through a gloo group: the NCCL unique id travels, and without a GPU
vers_comm_create fails loudly on every rank (no CPU fallback)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        import oracle as vo
        from gloo_model import chained_accumulate, exchange_rows_by_list, gather_init_centroids
        from vers_b200.sharded import balanced_list_owners, shard_bounds

        n, dim, C = 5003, 24, 13
        rows = vo.synth(1, n, dim, normalize=False)
        rows[17, 3] = -0.0  # must survive the centroid exchange bit for bit
        init = np.array([17, 4999, 17, 2500, 2501, 0, 5002, 1234, 1, 2, 3, 4, 4000], np.uint64)
        r0, nl = shard_bounds(n, rank, ws)
        local = torch.from_numpy(rows[r0:r0 + nl].copy())
        cents = gather_init_centroids(local, r0, nl, init, dim).numpy()
        assert np.array_equal(cents.view(np.uint32), rows[init.astype(np.int64)].view(np.uint32))
        assign = vo.assign(rows, cents)
        sums = torch.zeros((C, dim), dtype=torch.float32)
        counts = torch.zeros((C,), dtype=torch.int64)

        def step():
            s, c = sums.numpy(), counts.numpy().view(np.uint64)
            vo.partial_sums(rows[r0:r0 + nl], assign[r0:r0 + nl], C, s, c)

        chained_accumulate([sums, counts], step)
        new = vo.finalize_centroids(sums.numpy(), counts.numpy().view(np.uint64))
        want, wcnt = vo.update(rows, assign, C)  # the single-process reference order
        assert np.array_equal(counts.numpy().view(np.uint64), wcnt)
        assert np.array_equal(new.view(np.uint32), want.view(np.uint32)), "chained sums differ from global row order"
        # a plain all-reduce would NOT reproduce the reference association on this input
        part = np.zeros((C, dim), np.float32)
        pc = np.zeros(C, np.uint64)
        vo.partial_sums(rows[r0:r0 + nl], assign[r0:r0 + nl], C, part, pc)
        t = torch.from_numpy(part.copy())
        dist.all_reduce(t)
        ar = vo.finalize_centroids(t.numpy(), wcnt)
        sharded, _ = vo.update(rows, assign, C, shards=ws)
        assert np.array_equal(ar.view(np.uint32), sharded.view(np.uint32))  # == the oracle's sharded-order mode
        assert not np.array_equal(ar.view(np.uint32), want.view(np.uint32))

        # list-sharded index build: the all-to-all hands every rank whole lists, in ascending id order, balanced
        a_local = torch.from_numpy(assign[r0:r0 + nl].astype(np.int64))
        rr, rid, ras, owner = exchange_rows_by_list(local, a_local, r0, C, balanced_list_owners)
        owner = owner.numpy()
        sizes = np.bincount(assign.astype(np.int64), minlength=C)
        # the library's owner table (C++, vers_sharded_list_owners) == the specification: largest list first onto the
        # least-loaded rank, ties to the lowest rank / lowest list
        spec_order = np.lexsort((np.arange(C), -sizes))
        load, spec = np.zeros(ws, np.int64), np.zeros(C, np.int64)
        for c in spec_order:
            r = int(np.argmin(load))
            spec[c] = r
            load[r] += sizes[c]
        assert np.array_equal(owner, spec)
        mine = np.flatnonzero(owner[assign.astype(np.int64)] == rank)  # global ids this rank must end up with
        assert np.array_equal(rid.numpy(), mine), "received ids are not the owned lists' rows in ascending order"
        assert np.array_equal(ras.numpy().astype(np.int64), assign[mine].astype(np.int64))
        assert np.array_equal(rr.numpy().view(np.uint32), rows[mine].view(np.uint32))
        loads = np.array([sizes[owner == r].sum() for r in range(ws)])
        assert loads.max() - loads.min() <= sizes.max(), "largest-first placement keeps the ranks within one list"

        # bootstrap of the real Comm over this gloo group: without a CUDA device it must fail loudly, on every rank
        import vers_b200 as vb
        from vers_b200.sharded import Comm

        try:
            vb.Context(0)
            had_gpu = True
        except vb.VersError as e:
            had_gpu = False
            assert e.code == vb.ERR_CUDA
        if not had_gpu:
            class _NoCtx:  # the context handle is only dereferenced after the CUDA check
                h = None
            try:
                Comm(_NoCtx(), rank, ws, unique_id=bytes(128))
                raise AssertionError("vers_comm_create must fail without a context")
            except vb.VersError as e:
                assert e.code in (vb.ERR_ARG, vb.ERR_CUDA)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_chained_accumulate_world2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def test_shard_bounds_cover_all_rows():
    from vers_b200.sharded import shard_bounds

    for n in (0, 1, 7, 8, 10_000_000, 10_000_001):
        for ws in (1, 2, 3, 4, 8):
            got = [shard_bounds(n, r, ws) for r in range(ws)]
            assert sum(c for _, c in got) == n
            pos = 0
            for r0, c in got:
                assert r0 == pos or c == 0
                pos += c
