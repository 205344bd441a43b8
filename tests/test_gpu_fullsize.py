"""GPU parity at BASELINE.json's FULL sizes, through properties that do not need the CPU oracle to scan the whole
database: the tensor-core paths must agree bit for bit with the exact-order engine of the same library (which the
small-size tests pin to the oracle), results must not depend on scheduling (dynamic work items, per-query shared
bounds published at run time) nor on the position of a query in its batch."""
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


def test_c4_full_size_search_properties(vb, ctx):
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

    # a database row queried for itself is its own nearest neighbour at distance 0
    rows = ds.download(1234567, 8)
    ids5, d5, _ = idx.search_batch(rows, 1, nprobe=nprobe)
    # (8 queries: exact-order probe; the list scan still takes the tensor-core path)
    assert np.array_equal(ids5[:, 0], np.arange(1234567, 1234575, dtype=np.uint64)) and np.all(d5[:, 0] == 0.0)
    idx.close()
    ds.close()


def test_c5_full_size_assign_matches_exact_order_on_a_slice(vb, ctx):
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
