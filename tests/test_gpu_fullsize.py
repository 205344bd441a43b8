"""GPU parity at BASELINE.json's FULL sizes.

Two kinds of checks per config:
  * DIRECT oracle spot checks: the CPU oracle regenerates (include/vers_synth.h) exactly the rows a result depends on
    — the rows of the 32 probed inverted lists of a query (C4), a random row sample (C4/C5 assignments), the whole
    1M x 300 table (C2, C3) — and must return the same ids, distance bits, assignments and forest the device did.
  * size-independent properties: the tensor-core paths agree bit for bit with the exact-order engine of the same
    library, results do not depend on scheduling (dynamic work items, per-query shared bounds published at run time)
    nor on the position of a query in its batch."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _or_skip_nomem(fn):
    import vers_b200 as vb

    try:
        return fn()
    except vb.VersError as e:  # a smaller GPU: the full-size configs need ~70 GB
        if e.code == vb.ERR_NOMEM or "memory" in str(e).lower():
            pytest.skip(f"not enough device memory for the full-size config: {e}")
        raise


def test_c4_full_size_search_properties(vb, vo, ctx):
    """BASELINE configs[3]: 10M x 768, nlist 4096, nprobe 32, k 10"""
    n, dim, C, nq, k, nprobe = 10_000_000, 768, 4096, 256, 10, 32
    ds = _or_skip_nomem(lambda: vb.Dataset.synth(ctx, 1, n, dim, kind=1, n_centers=65536, center_seed=7, row0=0,
                                                 normalize=True))
    init = vb.synth_init_rows(3, 1, C, n)
    idx = _or_skip_nomem(lambda: vb.IVFFlatIndex.build_index(C, 1, 2, None, init_rows=init, ctx=ctx, dataset=ds))
    sizes = idx.list_sizes
    assert int(sizes.sum()) == n and sizes.shape == (C,)
    q = vb.Dataset.synth(ctx, 2, nq, dim, kind=1, n_centers=65536, center_seed=7, row0=0, normalize=True).download()

    ids0, d0, c0 = idx.search_batch(q, k, nprobe=nprobe)  # tensor-core probe + list scan, rerank, certificate
    st = idx.last_search_stats()
    assert st["reranked"] == nq * 32 and st["probe_reranked"] == nq * 64
    assert st["uncertified_queries"] <= nq // 16 and st["max_candidate_error"] < 5e-5
    assert np.all(c0 == k) and np.all(ids0 < n)
    assert np.all(np.diff(d0.astype(np.float64), axis=1) >= 0), "distances must come out ascending"
    for row in ids0:
        assert len(set(row.tolist())) == k, "an id must not appear twice in one result"

    # scheduling independence: same call again, bit-identical
    ids1, d1, _ = idx.search_batch(q, k, nprobe=nprobe)
    assert np.array_equal(ids0, ids1) and np.array_equal(bits(d0), bits(d1))

    # batch-position independence: reversed batch, and a sub-batch small enough for the exact-order probe
    ids2, d2, _ = idx.search_batch(q[::-1].copy(), k, nprobe=nprobe)
    assert np.array_equal(ids0, ids2[::-1]) and np.array_equal(bits(d0), bits(d2[::-1]))
    ids3, d3, _ = idx.search_batch(q[:17].copy(), k, nprobe=nprobe)
    assert np.array_equal(ids0[:17], ids3) and np.array_equal(bits(d0[:17]), bits(d3))

    # the exact-order engine (mode 1: no tensor cores anywhere) returns the same ids and distance bits
    idx.set_mode(1)
    try:
        ids4, d4, _ = idx.search_batch(q[:64].copy(), k, nprobe=nprobe)
    finally:
        idx.set_mode(0)
    assert np.array_equal(ids0[:64], ids4) and np.array_equal(bits(d0[:64]), bits(d4))

    # ---- DIRECT oracle spot check (ivfflat.rs:153-198 with the nprobe extension) on 16 queries spread over the batch:
    # the oracle ranks the 4096 downloaded centroids itself, regenerates the rows of the 32 lists it probes from the
    # synthetic-data specification (not from the device) and searches that sub-database; ids and distance bits must
    # equal what the device returned for the full 10M-row index.
    assign = idx.assignments  # [n] u64: final assign pass of build_kmeans (ivfflat.rs:96-98)
    cents = idx.centroids
    order = np.argsort(assign, kind="stable")  # ids of every list, ascending (ivfflat.rs:123-127)
    starts = np.searchsorted(assign[order], np.arange(C + 1))
    for qi in range(0, nq, nq // 16):
        dc = np.array([vo.l2sq(q[qi], cents[c]) for c in range(C)], np.float32)
        probed = np.lexsort((np.arange(C), dc))[:nprobe]  # stable sort by distance (ivfflat.rs:155-161)
        sub_ids = np.sort(np.concatenate([order[starts[c]:starts[c + 1]] for c in probed])).astype(np.uint64)
        sub_rows = vo.synth_rows(1, sub_ids, dim, kind=1, n_centers=65536, center_seed=7, normalize=True)
        off, lr = vo.ivf_lists(assign[sub_ids.astype(np.int64)], C)
        oi, od, oc = vo.ivf_search(sub_rows, cents, off, lr, q[qi:qi + 1], k, nprobe=nprobe)
        assert oc[0] == k
        assert np.array_equal(sub_ids[oi[0].astype(np.int64)], ids0[qi]), f"query {qi}: ids differ from the oracle"
        assert np.array_equal(bits(od[0]), bits(d0[qi])), f"query {qi}: distance bits differ from the oracle"
    # the assignments themselves: 2048 random rows re-assigned by the oracle from the final centroids
    rng = np.random.default_rng(5)
    pick = np.sort(rng.choice(n, 2048, replace=False)).astype(np.uint64)
    prow = vo.synth_rows(1, pick, dim, kind=1, n_centers=65536, center_seed=7, normalize=True)
    assert np.array_equal(vo.assign(prow, cents), assign[pick.astype(np.int64)]), "assignments differ from the oracle"
    del assign, order

    # a database row queried for itself is its own nearest neighbour at distance 0
    rows = ds.download(1234567, 8)
    ids5, d5, _ = idx.search_batch(rows, 1, nprobe=nprobe)
    # (8 queries: exact-order probe; the list scan still takes the tensor-core path)
    assert np.array_equal(ids5[:, 0], np.arange(1234567, 1234575, dtype=np.uint64)) and np.all(d5[:, 0] == 0.0)
    idx.close()
    ds.close()


def test_c5_full_size_assign_matches_exact_order_on_a_slice(vb, vo, ctx):
    """BASELINE configs[4]: 50M x 128, 16384 centroids — one assign pass on the tensor cores; the first 100k rows are
    re-assigned by the exact-order engine from the same centroids and must get the same clusters"""
    n, dim, C, m = 50_000_000, 128, 16384, 100_000
    ds = _or_skip_nomem(lambda: vb.Dataset.synth(ctx, 1, n, dim, kind=1, n_centers=65536, center_seed=7, row0=0,
                                                 normalize=False))
    init = vb.synth_init_rows(3, 1, C, n)[0]
    km = vb.KMeans(ds, C)
    km.init_from_rows(init)
    _or_skip_nomem(km.assign_step)
    flagged = km.last_uncertified_rows
    # first pass from random data rows as centroids (duplicates among the draws tie exactly): 0.5 % measured
    assert flagged < n // 50, "the certificate should fail for a small fraction of the rows only"
    a = km.assignments()
    assert a.shape == (n,) and int(a.max()) < C
    cents = km.centroids()

    # DIRECT oracle check (ivfflat.rs:29-46): 4096 random rows of the 50M, regenerated and assigned on the CPU
    rng = np.random.default_rng(6)
    pick = np.sort(rng.choice(n, 4096, replace=False)).astype(np.uint64)
    prow = vo.synth_rows(1, pick, dim, kind=1, n_centers=65536, center_seed=7, normalize=False)
    assert np.array_equal(vo.assign(prow, cents[:, :dim]), a[pick.astype(np.int64)]), "assignments differ from the oracle"

    ds2 = vb.Dataset.synth(ctx, 1, m, dim, kind=1, n_centers=65536, center_seed=7, row0=0, normalize=False)
    km2 = vb.KMeans(ds2, C)
    km2.set_mode(1)
    km2.set_centroids(cents)
    km2.assign_step()
    assert np.array_equal(km2.assignments(), a[:m])
    km2.close()
    km.close()
    ds2.close()
    ds.close()


@pytest.mark.parametrize("metric", [0, 1])
def test_c2_full_size_exhaustive_matches_oracle(vb, vo, ctx, metric):
    """BASELINE configs[1]: exhaustive top-10 over 1M x 300, 1000-query batch (utils.rs:68-82; cosine distance
    base.rs:155).  The oracle scans the whole table for 8 of the queries (both ends and the middle of the batch);
    single-query calls (the reference's own API shape) must return the same as the batch."""
    n, dim, nq, k = 1_000_000, 300, 1000, 10
    ds = vb.Dataset.synth(ctx, 1, n, dim, kind=1, n_centers=65536, center_seed=7, row0=0, normalize=True)
    q = vo.synth(2, nq, dim, kind=1, n_centers=65536, center_seed=7, normalize=True)
    ids, d, cnt = vb.search_exhaustive_batch(ds, q, k, metric)
    assert np.all(cnt == k)
    rows = vo.synth(1, n, dim, kind=1, n_centers=65536, center_seed=7, normalize=True)
    pick = [0, 1, 2, 499, 500, 777, 998, 999]
    oi, od, oc = vo.exhaustive(rows, q[pick], k, metric)
    assert np.array_equal(ids[pick], oi), "ids differ from the oracle"
    assert np.array_equal(bits(d[pick]), bits(od)), "distance bits differ from the oracle"
    for j, qi in enumerate(pick[:3]):  # nq = 1: the small-batch streaming kernel
        i1, d1, _ = vb.search_exhaustive_batch(ds, q[qi:qi + 1], k, metric)
        assert np.array_equal(i1[0], oi[j]) and np.array_equal(bits(d1[0]), bits(od[j]))
    i8, d8, _ = vb.search_exhaustive_batch(ds, q[pick], k, metric)  # nq = 8
    assert np.array_equal(i8, oi) and np.array_equal(bits(d8), bits(od))
    ds.close()


def test_c3_full_size_forest_and_search_match_oracle(vb, vo, ctx):
    """BASELINE configs[2]: hyperplane forest on 1M x 300, 16 trees, max_size 100 (lsh.rs:132-161, :264-282).  The
    oracle builds the same forest on the CPU: tree shapes, plane bits, leaf members and the search results of a
    64-query batch must be identical."""
    from test_gpu_parity import _same_forest

    n, dim, T, max_size, k = 1_000_000, 300, 16, 100, 10
    rows = vo.synth(1, n, dim, kind=1, n_centers=65536, center_seed=7, normalize=True)
    q = vo.synth(2, 64, dim, kind=1, n_centers=65536, center_seed=7, normalize=True)
    g = vb.ANNIndex.build_index(T, max_size, rows, None, seed=4, ctx=ctx)
    o = vo.LSH(rows, None, T, max_size, 4)
    assert g.info()["num_values"] == o.num_values
    _same_forest(g, o, T)
    ids, d, cnt = g.search_batch(q, k)
    oi, od, oc = o.search(q, k)
    assert np.array_equal(cnt, oc) and np.array_equal(ids, oi) and np.array_equal(bits(d), bits(od))
    g.close()
