"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs.
Bar: bit-exact ids / assignments / hash bits AND bit-exact distances / centroids (the device computes in the
reference's own summation order, so the 1e-5 tolerance of the north star is met with zero slack)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def data(vo, n, dim, seed=1, kind=1, n_centers=16, normalize=True):
    return vo.synth(seed, n, dim, kind=kind, n_centers=n_centers, center_seed=7, normalize=normalize)


# ---------------------------------------------------------------------------------------------- datasets
@pytest.mark.parametrize("dim", [300, 128, 7, 33])
@pytest.mark.parametrize("kind", [0, 1])
def test_synth_and_normalize_bits(vb, vo, ctx, dim, kind):
    n = 1000
    ds = vb.Dataset.synth(ctx, 11, n, dim, kind=kind, n_centers=5, center_seed=9, row0=123, normalize=True)
    want = vo.synth(11, n, dim, kind=kind, n_centers=5, center_seed=9, row0=123, normalize=True)
    got = ds.download()
    assert np.array_equal(bits(got), bits(want))
    assert ds.id_base == 123


def test_upload_strided_like_reference_vector(vb, vo, ctx):
    # size_of::<Vector<300>>() = 1280 B => stride 320 floats (base.rs:15)
    rows = data(vo, 257, 300)
    buf = np.zeros((257, 320), np.float32)
    buf[:, :300] = rows
    buf[:, 300:] = 777.0  # padding garbage must be ignored
    ds = vb.Dataset.upload_strided(ctx, buf, 257, 300, 320)
    assert np.array_equal(bits(ds.download()), bits(rows))


def test_normalize_tiny_rows_unchanged(vb, vo, ctx):
    rows = data(vo, 64, 20, normalize=False)
    rows[3] = 0.0
    rows[5] = 1e-9
    ds = vb.Dataset.upload(ctx, rows)
    ds.normalize()
    assert np.array_equal(bits(ds.download()), bits(vo.normalize_rows(rows)))


# ---------------------------------------------------------------------------------------------- exhaustive
@pytest.mark.parametrize("n,dim,nq,k", [(10000, 300, 100, 10), (5000, 128, 3, 1), (777, 33, 40, 128), (50, 7, 9, 10),
                                        (20000, 768, 33, 10)])
@pytest.mark.parametrize("metric", [0, 1])
def test_flat_search_matches_search_exhaustive(vb, vo, ctx, n, dim, nq, k, metric):
    rows = data(vo, n, dim)
    q = data(vo, nq, dim, seed=2)
    ds = vb.Dataset.upload(ctx, rows, id_base=1000)
    ids, d, cnt = vb.search_exhaustive_batch(ds, q, k, metric)
    oi, od, oc = vo.exhaustive(rows, q, k, metric, id_base=1000)
    assert np.array_equal(cnt, oc)
    assert np.array_equal(ids, oi)
    assert np.array_equal(bits(d), bits(od))


@pytest.mark.parametrize("n,dim,nq,k", [(40000, 300, 1, 10), (40000, 300, 8, 10), (9999, 768, 1, 10), (9999, 768, 5, 3),
                                        (30001, 128, 2, 128), (4097, 36, 7, 33), (31, 300, 1, 10), (1000, 7, 4, 1),
                                        (33, 20, 6, 64)])
@pytest.mark.parametrize("metric", [0, 1])
def test_flat_search_small_batch_streaming_kernel(vb, vo, ctx, n, dim, nq, k, metric):
    """nq <= 8: flat_stream_kernel (bulk-copy staged row tiles, lane = row, one exact-order chain per query): odd and
    even float4 row strides (staggered lanes), a last partial tile, tables smaller than a tile, k up to 128, both
    metrics; the exact-order tile engine (flat mode 1) must agree as well"""
    rows = data(vo, n, dim)
    rows[n // 2] = rows[3]  # a duplicate row: equal distances, the lower id first
    q = data(vo, nq, dim, seed=2)
    q[0] = rows[3]
    ds = vb.Dataset.upload(ctx, rows, id_base=77)
    ids, d, cnt = vb.search_exhaustive_batch(ds, q, k, metric)
    oi, od, oc = vo.exhaustive(rows, q, k, metric, id_base=77)
    assert np.array_equal(cnt, oc) and np.array_equal(ids, oi) and np.array_equal(bits(d), bits(od))
    ds.set_flat_mode(1)
    ids1, d1, cnt1 = vb.search_exhaustive_batch(ds, q, k, metric)
    assert np.array_equal(ids1, oi) and np.array_equal(bits(d1), bits(od)) and np.array_equal(cnt1, oc)


def test_flat_search_ties_broken_by_id(vb, vo, ctx):
    rows = data(vo, 300, 64)
    rows[100:200] = rows[7]  # 101 identical rows -> identical distances
    q = data(vo, 5, 64, seed=2)
    q[0] = rows[7]
    ds = vb.Dataset.upload(ctx, rows)
    ids, d, cnt = vb.search_exhaustive_batch(ds, q, 20)
    oi, od, oc = vo.exhaustive(rows, q, 20)
    assert np.array_equal(ids, oi) and np.array_equal(bits(d), bits(od))
    assert list(ids[0][:3]) == [7, 100, 101]


def test_flat_search_fewer_rows_than_k_and_single(vb, vo, ctx):
    rows = data(vo, 6, 300)
    q = data(vo, 2, 300, seed=2)
    ds = vb.Dataset.upload(ctx, rows)
    ids, d, cnt = vb.search_exhaustive_batch(ds, q, 10)
    oi, od, oc = vo.exhaustive(rows, q, 10)
    assert list(cnt) == [6, 6]
    assert np.array_equal(ids, oi)
    got = vb.search_exhaustive(ds, q[0], 3)
    assert [g[0] for g in got] == list(oi[0][:3])


# ---------------------------------------------------------------------------------------------- k-means
@pytest.mark.parametrize("n,dim,C", [(10000, 300, 16), (4000, 128, 100), (3000, 768, 5), (999, 33, 64)])
def test_assign_to_clusters_bit_exact(vb, vo, ctx, n, dim, C):
    rows = data(vo, n, dim)
    cents = rows[vo.init_rows(3, 1, C, n)[0].astype(np.int64)].copy()
    ds = vb.Dataset.upload(ctx, rows)
    got = vb.assign_to_clusters(ds, cents)
    assert np.array_equal(got, vo.assign(rows, cents))


def test_assign_first_minimum_on_duplicate_centroids(vb, vo, ctx):
    rows = data(vo, 2000, 64)
    cents = rows[[5, 9, 5, 9, 11, 5]].copy()  # duplicates: the FIRST minimum must win (min_by, ivfflat.rs:36-42)
    ds = vb.Dataset.upload(ctx, rows)
    got = vb.assign_to_clusters(ds, cents)
    want = vo.assign(rows, cents)
    assert np.array_equal(got, want)
    assert not np.any(np.isin(got, [2, 3, 5]))
    # the split-precision candidate pass (mode 2) keeps (best, second best): it sees exact ties between the twins, so
    # every row whose nearest centroid has a twin must fail its certificate and be redone in exact order.  The
    # tf32-first pass (mode 0, ld <= 128) re-ranks its four best candidates in exact order inside the kernel: twins and
    # triplets are resolved there (first minimum wins) without a redo.
    for mode in (0, 2, 3):
        km = vb.KMeans(ds, 6)
        km.set_mode(mode)
        km.set_centroids(cents)
        km.assign_step()
        assert np.array_equal(km.assignments(), want)
        if mode == 2:
            assert km.last_uncertified_rows >= int((want != 4).sum()) > 0
        km.close()


@pytest.mark.parametrize("n,dim,C,normalize", [(5000, 128, 1000, False), (3001, 100, 65, True), (777, 36, 3, True),
                                               (2048, 64, 1, True), (1500, 128, 5, False), (513, 32, 200, True)])
@pytest.mark.parametrize("km_mode", [0, 3])
def test_assign_tf32_first_kernel_bit_exact(vb, vo, ctx, n, dim, C, normalize, km_mode):
    """tc_assign1_kernel (ld <= 128): one MMA per K step (mode 0: kind::f16 with both operands scaled into fp16's range,
    mode 3: kind::tf32), top-4 + exact rerank + certificate; tile edges (C not a multiple of 64, fewer than five
    centroids, a last row-block pair that is half empty, an odd number of 32-float chunks), duplicated centroids"""
    rows = data(vo, n, dim, normalize=normalize)
    pick = vo.init_rows(5, 1, C, n)[0].astype(np.int64)
    cents = rows[pick].copy()
    if C >= 8:
        cents[C // 2] = cents[0]  # an exact twin
        cents[1] = 0.5 * (cents[2] + cents[3])  # and a centroid that is not a data row
    ds = vb.Dataset.upload(ctx, rows)
    km = vb.KMeans(ds, C)
    km.set_mode(km_mode)
    km.set_centroids(cents)
    km.assign_step()
    assert np.array_equal(km.assignments(), vo.assign(rows, cents))
    assert km.last_uncertified_rows <= n // 4
    km.close()


@pytest.mark.parametrize("km_mode", [0, 3, 2])
def test_assign_many_row_blocks_per_cta(vb, vo, ctx, km_mode):
    """enough rows that every persistent CTA walks several row blocks (148 CTAs x 256 rows x 4), enough centroids for 20
    tiles: the hand-over between row blocks (A operand in tensor memory, norm ring, accumulator buffers) and the
    reconvergence of the warps after their divergent selection only show after the first block — a regression that
    passed every small-shape test returned wrong clusters for 0.2 % of the rows from the second block on"""
    n, dim, C = 160_000, 128, 1250
    rows = data(vo, n, dim, n_centers=700, normalize=False)
    pick = vo.init_rows(5, 1, C, n)[0].astype(np.int64)
    cents = rows[pick].copy()
    ds = vb.Dataset.upload(ctx, rows)
    ref = vb.KMeans(ds, C)
    ref.set_mode(1)  # exact-order engine
    ref.set_centroids(cents)
    ref.assign_step()
    want = ref.assignments()
    ref.close()
    sample = np.arange(0, n, 37)
    assert np.array_equal(want[sample], vo.assign(rows[sample], cents))  # the engine itself against the oracle
    km = vb.KMeans(ds, C)
    km.set_mode(km_mode)
    km.set_centroids(cents)
    for rep in range(2):
        km.assign_step()
        got = km.assignments()
        assert np.array_equal(got, want), f"{int((got != want).sum())} rows differ, first {np.flatnonzero(got != want)[:5]}"
        assert km.last_uncertified_rows <= n // 20
    km.close()


def test_assign_f16_kernel_scaling_and_overflow(vb, vo, ctx):
    """kind::f16 assign (mode 0, ld <= 128): rows of very different magnitudes (1e-4 .. 3e4, far outside fp16's own range: the
    kernel scales by a power of two chosen from max ||row||), and centroids set far beyond any row (their fp16 image
    overflows): the assignments must still be the oracle's — the second case through the exact redo of every row"""
    n, dim, C = 3000, 96, 40
    rows = data(vo, n, dim, normalize=False)
    rows *= np.float32(3.0e4)
    rows[::5] *= np.float32(1.0e-8)
    pick = vo.init_rows(5, 1, C, n)[0].astype(np.int64)
    cents = rows[pick].copy()
    ds = vb.Dataset.upload(ctx, rows)
    km = vb.KMeans(ds, C)
    km.set_mode(0)
    km.set_centroids(cents)
    km.assign_step()
    assert np.array_equal(km.assignments(), vo.assign(rows, cents))
    assert km.last_uncertified_rows <= n // 2
    far = cents.copy()
    far[3] *= np.float32(1.0e6)  # not a mean of rows: does not fit the rows' scale
    km.set_centroids(far)
    km.assign_step()
    assert np.array_equal(km.assignments(), vo.assign(rows, far))
    assert km.last_uncertified_rows == n
    km.close()


def test_assign_nan_row_panics_like_the_reference(vb, vo, ctx):
    """partial_cmp(..).unwrap() on a NaN distance panics (ivfflat.rs:39): both assign paths must report it instead of
    writing an out-of-range cluster"""
    rows = data(vo, 1000, 64)
    rows[123, 7] = np.nan
    cents = rows[[1, 2, 3, 4, 5, 6, 7, 8]].copy()
    ds = vb.Dataset.upload(ctx, rows)
    for mode in (0, 1, 2, 3):
        km = vb.KMeans(ds, 8)
        km.set_mode(mode)
        km.set_centroids(cents)
        with pytest.raises(vb.VersPanic):
            km.assign_step()
        km.close()


@pytest.mark.parametrize("n,dim,C", [(10000, 300, 16), (5000, 128, 64), (1000, 20, 7)])
def test_update_centroids_bit_exact(vb, vo, ctx, n, dim, C):
    rows = data(vo, n, dim, normalize=False)
    rng = np.random.default_rng(5)
    assign = rng.integers(0, C, n).astype(np.uint64)
    assign[assign == 3] = 2  # cluster 3 empty -> zero vector (ivfflat.rs:63-67)
    ds = vb.Dataset.upload(ctx, rows)
    cents, counts = vb.update_centroids(ds, assign, C)
    oc, on = vo.update(rows, assign, C)
    assert np.array_equal(counts, on)
    assert np.array_equal(bits(cents), bits(oc))
    assert counts[3] == 0 and not cents[3].any()


@pytest.mark.parametrize("km_mode", [0, 1, 2, 3])
@pytest.mark.parametrize("n,dim,C", [(10000, 300, 16), (6000, 128, 300), (4097, 36, 129)])
def test_kmeans_fit_and_cost_bit_exact(vb, vo, ctx, km_mode, n, dim, C):
    """km_mode 0: tensor-core candidate argmin + certificate + exact redo of uncertified rows (tf32-first kernel for
    ld <= 128, split-precision kernel otherwise); km_mode 1: exact order only; km_mode 2: split-precision kernel.  Iterations, assignments, centroid bits and cost bits must equal the oracle's."""
    rows = data(vo, n, dim)
    init = vo.init_rows(3, 1, C, n)[0]
    ds = vb.Dataset.upload(ctx, rows)
    km = vb.KMeans(ds, C)
    km.set_mode(km_mode)
    km.init_from_rows(init)
    iters = km.fit(20)
    cents, assign, oit = vo.kmeans_fit(rows, init, 20)
    assert iters == oit
    assert np.array_equal(km.assignments(), assign)
    assert np.array_equal(bits(km.centroids()), bits(cents))
    cost = km.cost_step(0.0)
    assert bits(np.float32(cost)) == bits(vo.kmeans_cost(rows, cents, assign))


def test_kmeans_chained_shards_reproduce_global_order(vb, vo, ctx):
    """two row shards whose sums are chained in row order == the single-GPU / reference association"""
    import torch

    n, dim, C = 6000, 128, 32
    rows = data(vo, n, dim, normalize=False)
    cents0 = rows[vo.init_rows(3, 1, C, n)[0].astype(np.int64)].copy()
    assign = vo.assign(rows, cents0)
    want, wcnt = vo.update(rows, assign, C)
    h = n // 2 + 17
    shards = [vb.Dataset.upload(ctx, rows[:h]), vb.Dataset.upload(ctx, rows[h:], id_base=h)]
    kms = []
    for ds in shards:
        km = vb.KMeans(ds, C)
        km.set_centroids(cents0)
        km.assign_step()
        kms.append(km)
    ld = shards[0].ld
    sums = torch.zeros(C * ld, dtype=torch.float32, device="cuda")
    counts = torch.zeros(C, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    for km in kms:  # rank order == row order
        km.sums_step_dev(sums.data_ptr(), counts.data_ptr())
    ctx.sync()
    changed = kms[0].finalize_step_dev(sums.data_ptr(), counts.data_ptr())
    assert changed
    assert np.array_equal(counts.cpu().numpy().astype(np.uint64), wcnt)
    assert np.array_equal(bits(kms[0].centroids()), bits(want))


@pytest.mark.parametrize("n,dim,nq,k", [(50000, 300, 100, 10), (20000, 768, 33, 10), (70000, 128, 257, 1),
                                        (6000, 96, 64, 37), (150000, 64, 40, 64), (30000, 300, 20, 10),
                                        (9000, 100, 9, 16), (200000, 48, 32, 3)])
def test_flat_search_tensor_core_path_bit_exact(vb, vo, ctx, n, dim, nq, k):
    """large batches take tensor-core candidate keys -> exact rerank -> certificate (-> exact redo): same ids and
    distance bits as search_exhaustive, and the same as the exact-order engine (mode 1)"""
    rows = data(vo, n, dim, n_centers=64)
    q = data(vo, nq, dim, seed=2, n_centers=64)
    ds = vb.Dataset.upload(ctx, rows, id_base=77)
    ids, d, cnt = vb.search_exhaustive_batch(ds, q, k, 0)
    st = ds.last_flat_search_stats()
    oi, od, oc = vo.exhaustive(rows, q, k, 0, id_base=77)
    assert np.array_equal(cnt, oc) and np.array_equal(ids, oi) and np.array_equal(bits(d), bits(od))
    # batches of >= 96 queries over rows of <= 320 floats with k <= 16 run tc_flat_kernel (ONE tf32 MMA per K step:
    # candidate values carry the tf32 rounding of both operands); the others (from 9 queries on) the split-precision
    # list-scan kernel
    single_tf32 = nq >= 96 and dim <= 320 and k <= 16 and n >= 4096
    assert st["reranked"] > 0 and st["uncertified_queries"] <= nq // 4
    assert st["max_candidate_error"] < (2e-3 if single_tf32 else 2e-5)
    ds.set_flat_mode(2)  # the list-scan style kernel for every eligible batch
    ids2, d2, cnt2 = vb.search_exhaustive_batch(ds, q, k, 0)
    assert np.array_equal(ids2, oi) and np.array_equal(bits(d2), bits(od))
    assert ds.last_flat_search_stats()["max_candidate_error"] < 2e-5
    ds.set_flat_mode(1)
    ids1, d1, cnt1 = vb.search_exhaustive_batch(ds, q, k, 0)
    assert np.array_equal(ids1, oi) and np.array_equal(bits(d1), bits(od))
    assert ds.last_flat_search_stats()["reranked"] == 0


@pytest.mark.parametrize("n,dim,nq,k,normalize", [(100000, 300, 1000, 10, True), (40000, 128, 130, 16, True),
                                                  (5000, 64, 96, 5, False), (9000, 36, 200, 1, True),
                                                  (4100, 320, 128, 10, False), (33333, 100, 257, 10, True),
                                                  (9000, 300, 40, 10, True), (12000, 64, 33, 5, False),
                                                  (20000, 128, 95, 16, True)])
def test_flat_search_query_block_kernel_bit_exact(vb, vo, ctx, n, dim, nq, k, normalize):
    """tc_flat_kernel: 128 queries resident in tensor memory, the table streamed in 64-row tiles once per query block;
    last query block partly empty, last slice / last tile ragged, K chunks 2..10, unnormalised rows; batches below 96
    queries take the list-scan style kernel with 32 candidates and the shared bound (same statistics)"""
    rows = data(vo, n, dim, n_centers=50, normalize=normalize)
    q = data(vo, nq, dim, seed=2, n_centers=50, normalize=normalize)
    q[3] = rows[n - 1]  # the very last row of the table is its own nearest neighbour
    ds = vb.Dataset.upload(ctx, rows, id_base=5)
    ids, d, cnt = vb.search_exhaustive_batch(ds, q, k, 0)
    st = ds.last_flat_search_stats()
    oi, od, oc = vo.exhaustive(rows, q, k, 0, id_base=5)
    assert np.array_equal(cnt, oc) and np.array_equal(ids, oi) and np.array_equal(bits(d), bits(od))
    assert ids[3][0] == n - 1 + 5 and d[3][0] == 0.0
    assert st["reranked"] == 32 * nq and st["uncertified_queries"] <= max(2, nq // 8)
    ids_b, d_b, _ = vb.search_exhaustive_batch(ds, q, k, 0)  # run-to-run: the shared bounds must not change anything
    assert np.array_equal(ids_b, ids) and np.array_equal(bits(d_b), bits(d))


def test_flat_search_query_block_kernel_ties_fall_back(vb, vo, ctx):
    """200 copies of one row: more ties than the 32 candidates, the certificate must fail and the exact redo order
    them by id (the other queries stay certified)"""
    rows = data(vo, 30000, 96, n_centers=32)
    rows[5000:5200] = rows[11]
    q = data(vo, 128, 96, seed=2, n_centers=32)
    q[0] = rows[11]
    ds = vb.Dataset.upload(ctx, rows)
    ids, d, cnt = vb.search_exhaustive_batch(ds, q, 10, 0)
    oi, od, oc = vo.exhaustive(rows, q, 10, 0)
    assert np.array_equal(ids, oi) and np.array_equal(bits(d), bits(od))
    assert list(ids[0]) == [11] + list(range(5000, 5009))
    assert 1 <= ds.last_flat_search_stats()["uncertified_queries"] <= 16


def test_flat_search_tensor_core_path_ties_fall_back(vb, vo, ctx):
    """200 copies of one row: more ties than candidates, the certificate must fail and the exact redo order them by id"""
    rows = data(vo, 30000, 96, n_centers=32)
    rows[5000:5200] = rows[11]
    q = data(vo, 48, 96, seed=2, n_centers=32)
    q[0] = rows[11]
    ds = vb.Dataset.upload(ctx, rows)
    ids, d, cnt = vb.search_exhaustive_batch(ds, q, 10, 0)
    oi, od, oc = vo.exhaustive(rows, q, 10, 0)
    assert np.array_equal(ids, oi) and np.array_equal(bits(d), bits(od))
    assert list(ids[0]) == [11] + list(range(5000, 5009))
    assert ds.last_flat_search_stats()["uncertified_queries"] >= 1


# ---------------------------------------------------------------------------------------------- HNSW distance offload
@pytest.mark.parametrize("dim", [3, 64, 67, 128, 300, 768, 2100])
def test_pair_distances_simd_bit_exact(vb, vo, ctx, dim):
    """vers_pair_distances_simd == Vector::cosine_similarity_simd / squared_euclidean_simd (base.rs:158-294) for a batch
    of (query, row id) pairs: 64-wide chunks, 4-wide chunks, scalar tail, ordered chunk sums; global ids with an id
    base; a missing row panics like id_to_vec.get(..).unwrap() (hnsw.rs:133)"""
    n, nq, npairs = 3000, 7, 5000
    rows = data(vo, n, dim)
    q = data(vo, nq, dim, seed=2)
    rng = np.random.default_rng(dim)
    pr = rng.integers(0, n, npairs).astype(np.uint64)
    pq = rng.integers(0, nq, npairs).astype(np.uint32)
    ds = vb.Dataset.upload(ctx, rows, id_base=1000)
    golden = np.load(os.path.join(os.path.dirname(__file__), "golden", "simd_small.npz"))
    for metric in (0, 1):
        got = vb.pair_distances_simd(ds, q, pr + 1000, pq, metric)
        assert np.array_equal(bits(got), bits(vo.pair_distances_simd(rows, q, pr, pq, metric)))
        one = vb.pair_distances_simd(ds, q[3], pr[:100] + 1000, None, metric)  # one node against its neighbours
        assert np.array_equal(bits(one), bits(vo.pair_distances_simd(rows, q[3:4], pr[:100], None, metric)))
        if dim in (300, 67):  # the committed golden vectors, straight through the C ABI
            grows = vo.synth(1, 500, dim, kind=1, n_centers=8, center_seed=7, normalize=True)
            gq = vo.synth(2, 6, dim, kind=1, n_centers=8, center_seed=7, normalize=True)
            gds = vb.Dataset.upload(ctx, grows)
            d = vb.pair_distances_simd(gds, gq, golden[f"pair_row_{dim}"], golden[f"pair_query_{dim}"], metric)
            assert np.array_equal(bits(d), golden[f"d_{dim}_m{metric}_bits"])
    assert vb.pair_distances_simd(ds, q, np.zeros(0, np.uint64), None, 1).shape == (0,)
    for bad_row, bad_q in ((999, 0), (1000 + n, 0), (1000, nq)):
        with pytest.raises(vb.VersPanic):
            vb.pair_distances_simd(ds, q, np.array([bad_row], np.uint64), np.array([bad_q], np.uint32), 1)


# ---------------------------------------------------------------------------------------------- IVFFlat
@pytest.fixture(scope="module")
def ivf_c1(vb, vo, ctx):
    """BASELINE config 1: 10k x 300 normalized, 16 clusters"""
    n, dim, C = 10000, 300, 16
    rows = data(vo, n, dim)
    init = vo.init_rows(3, 2, C, n)
    idx = vb.IVFFlatIndex.build_index(C, 2, 20, rows, init_rows=init, ctx=ctx)
    cents, assign, cost, best = vo.ivf_build_index(rows, C, 2, 20, init)
    return dict(rows=rows, idx=idx, cents=cents, assign=assign, cost=cost, best=best, C=C)


def test_ivf_build_index_bit_exact(ivf_c1):
    s = ivf_c1
    assert s["idx"].best_attempt == s["best"]
    assert bits(np.float32(s["idx"].best_cost)) == bits(np.float32(s["cost"]))
    assert np.array_equal(s["idx"].assignments, s["assign"])
    assert np.array_equal(bits(s["idx"].centroids), bits(s["cents"]))
    assert np.array_equal(s["idx"].list_sizes, np.bincount(s["assign"].astype(np.int64), minlength=s["C"]))


@pytest.mark.parametrize("exact_mode", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("k", [1, 10, 37, 100])
@pytest.mark.parametrize("nprobe", [0, 1, 4, 16])
def test_ivf_search_bit_exact(vo, ivf_c1, k, nprobe, exact_mode):
    """mode 0: tensor-core (TMA + tcgen05 TF32) candidate pass + exact-order rerank + certificate (+ exact redo when
    uncertified); mode 2: the same with an fp32 FMA candidate pass; mode 1: exact order everywhere; mode 4: candidates
    from the fp16 copy of the lists (kind::f16 MMA), exact fp32 rerank, Cauchy-Schwarz certificate.
    Every mode must return the oracle's ids AND distance bits."""
    s = ivf_c1
    q = data(vo, 100, 300, seed=2)
    off, lr = vo.ivf_lists(s["assign"], s["C"])
    s["idx"].set_mode(exact_mode)
    try:
        ids, d, cnt = s["idx"].search_batch(q, k, nprobe=nprobe)
        st = s["idx"].last_search_stats()
    finally:
        s["idx"].set_mode(0)
    oi, od, oc = vo.ivf_search(s["rows"], s["cents"], off, lr, q, k, nprobe=nprobe)
    assert np.array_equal(cnt, oc)
    assert np.array_equal(ids, oi)
    assert np.array_equal(bits(d), bits(od))
    if exact_mode != 1 and nprobe > 0 and k <= 16:
        assert st["reranked"] > 0  # the candidate path really ran
        # plain TF32 (3) and the 16-bit copy (4) certify less often on this small, dense set
        assert st["uncertified_queries"] <= (10 if exact_mode in (0, 2) else 100)
        assert st["max_candidate_error"] < {3: 3e-3, 4: 5e-4}.get(exact_mode, 2e-5)


@pytest.fixture(scope="module")
def ivf_wide(vb, vo, ctx):
    """enough lists (512) and queries (>= 32) for the tensor-core centroid probe"""
    n, dim, C = 24000, 96, 512
    rows = data(vo, n, dim, n_centers=300)
    init = vo.init_rows(3, 1, C, n)
    idx = vb.IVFFlatIndex.build_index(C, 1, 4, rows, init_rows=init, ctx=ctx)
    cents, assign, _, _ = vo.ivf_build_index(rows, C, 1, 4, init)
    assert np.array_equal(idx.assignments, assign)
    return dict(rows=rows, idx=idx, cents=cents, assign=assign, C=C, dim=dim)


@pytest.mark.parametrize("nprobe", [1, 8, 32, 64])
@pytest.mark.parametrize("nq", [32, 200, 333])
def test_ivf_tensor_core_probe_bit_exact(vo, ivf_wide, nq, nprobe):
    """centroid probe on the tensor cores (top-M keys -> exact-order rerank -> certificate -> exact redo of the
    uncertified queries): the probed lists, hence ids and distance bits, must equal the oracle's"""
    s = ivf_wide
    q = data(vo, nq, s["dim"], seed=2, n_centers=300)
    off, lr = vo.ivf_lists(s["assign"], s["C"])
    ids, d, cnt = s["idx"].search_batch(q, 10, nprobe=nprobe)
    st = s["idx"].last_search_stats()
    oi, od, oc = vo.ivf_search(s["rows"], s["cents"], off, lr, q, 10, nprobe=nprobe)
    assert np.array_equal(cnt, oc)
    assert np.array_equal(ids, oi)
    assert np.array_equal(bits(d), bits(od))
    assert st["probe_reranked"] > 0  # the tensor-core probe really ran
    assert st["uncertified_probe_queries"] <= nq // 8


def test_ivf_tensor_core_probe_large_centroid_table(vb, vo, ctx):
    """more than 8192 lists: the probe keeps per-chunk top-32 lists in the scan kernel (no dense key matrix)"""
    n, dim, C = 30000, 32, 9000
    rows = data(vo, n, dim, n_centers=500)
    cents = rows[::3][:C].copy()  # 9000 distinct rows as centroids
    assign = vo.assign(rows, cents).astype(np.uint64)
    idx = vb.IVFFlatIndex.from_parts(rows, cents, assign, ctx=ctx)
    q = data(vo, 96, dim, seed=2, n_centers=500)
    off, lr = vo.ivf_lists(assign, C)
    for nprobe in (4, 32):
        ids, d, cnt = idx.search_batch(q, 10, nprobe=nprobe)
        st = idx.last_search_stats()
        oi, od, oc = vo.ivf_search(rows, cents, off, lr, q, 10, nprobe=nprobe)
        assert np.array_equal(cnt, oc) and np.array_equal(ids, oi) and np.array_equal(bits(d), bits(od))
        assert st["probe_reranked"] > 0


def test_ivf_tensor_core_probe_falls_back_on_tied_centroids(vb, vo, ctx, ivf_wide):
    """100 identical centroids (more than the candidate list holds) are the nearest ones: the certificate cannot
    separate them, the exact redo must open the lowest-numbered ones like the reference's stable sort (ivfflat.rs:160)"""
    s = ivf_wide
    cents = s["cents"].copy()
    big = int(np.argmax(np.bincount(s["assign"].astype(np.int64), minlength=s["C"])))
    tied = np.array([c for c in range(100, 201) if c != big][:100])
    cents[tied] = cents[big]
    assign = vo.assign(s["rows"], cents).astype(np.uint64)
    members = np.flatnonzero(assign == min(big, int(tied[0])))  # first minimum: the lowest-numbered copy gets the rows
    assert members.size >= 60
    assign[members] = tied[np.arange(members.size) % 100]  # spread those rows over the tied lists
    idx = vb.IVFFlatIndex.from_parts(s["rows"], cents, assign, ctx=ctx)
    q = data(vo, 64, s["dim"], seed=2, n_centers=300)
    q[:8] = s["rows"][members[:8]]  # queries whose nearest centroids are the tied ones
    off, lr = vo.ivf_lists(assign, s["C"])
    for nprobe in (8, 32):
        ids, d, cnt = idx.search_batch(q, 10, nprobe=nprobe)
        st = idx.last_search_stats()
        oi, od, oc = vo.ivf_search(s["rows"], cents, off, lr, q, 10, nprobe=nprobe)
        assert np.array_equal(cnt, oc) and np.array_equal(ids, oi) and np.array_equal(bits(d), bits(od))
        assert st["uncertified_probe_queries"] >= 8


def test_ivf_candidate_path_falls_back_on_ties(vb, vo, ctx):
    """more exact duplicates than the candidate list holds: the rounding-error certificate cannot separate them,
    the queries must be redone in exact order and still return the reference's (distance, id) order"""
    n, dim, C, k = 6000, 96, 8, 10
    rows = data(vo, n, dim)
    rows[1000:1100] = rows[17]           # 101 copies of row 17
    rows[2000:2040] = rows[23]           # 41 copies of row 23
    init = vo.init_rows(3, 1, C, n)
    idx = vb.IVFFlatIndex.build_index(C, 1, 6, rows, init_rows=init, ctx=ctx)
    cents, assign, _, _ = vo.ivf_build_index(rows, C, 1, 6, init)
    off, lr = vo.ivf_lists(assign, C)
    q = data(vo, 40, dim, seed=2)
    q[0] = rows[17]
    q[1] = rows[23]
    q[2] = (rows[17] + np.float32(1e-4)).astype(np.float32)
    ids, d, cnt = idx.search_batch(q, k, nprobe=4)
    st = idx.last_search_stats()
    oi, od, oc = vo.ivf_search(rows, cents, off, lr, q, k, nprobe=4)
    assert np.array_equal(ids, oi) and np.array_equal(bits(d), bits(od)) and np.array_equal(cnt, oc)
    assert list(ids[0]) == [17] + list(range(1000, 1009))
    assert st["uncertified_queries"] >= 2  # the duplicate-heavy queries took the exact redo


@pytest.mark.parametrize("dim", [96, 300, 768])
def test_ivf_h16_candidate_copy_bit_exact(vb, vo, ctx, dim):
    """mode 4 (BASELINE configs[3] "16-bit candidate / fp32 rerank"): the scan streams an fp16 copy of the lists; ids and
    distance bits must still be the oracle's, most queries must certify (the copy's error bound is ~2e-4 on unit rows),
    and the copy must follow add (in place), a row too large for the copy's scale (rebuild) and add_batch (rebuild)."""
    n, C, k = 40000, 64, 10
    rows = data(vo, n, dim, n_centers=300)
    init = vo.init_rows(3, 1, C, n)
    idx = vb.IVFFlatIndex.build_index(C, 1, 4, rows, init_rows=init, ctx=ctx)
    cents, assign, _, _ = vo.ivf_build_index(rows, C, 1, 4, init)
    assert np.array_equal(idx.assignments, assign)
    idx.set_mode(4)
    q = data(vo, 200, dim, seed=2, n_centers=300)
    off, lr = vo.ivf_lists(assign, C)
    for nprobe in (1, 8, 32):
        ids, d, cnt = idx.search_batch(q, k, nprobe=nprobe)
        st = idx.last_search_stats()
        oi, od, oc = vo.ivf_search(rows, cents, off, lr, q, k, nprobe=nprobe)
        assert np.array_equal(cnt, oc) and np.array_equal(ids, oi) and np.array_equal(bits(d), bits(od))
        assert st["reranked"] > 0
        assert st["uncertified_queries"] <= 60
        assert st["max_candidate_error"] < 5e-4
    # single adds keep the copy current in place; the last one is 50x larger than anything the copy's scale admits
    extra = data(vo, 40, dim, seed=5, n_centers=300)
    extra[-1] *= np.float32(50.0)
    all_rows, all_assign = np.vstack([rows, extra]), list(assign)
    for i in range(extra.shape[0]):
        if i == extra.shape[0] - 1:  # search once more before the oversized row: the in-place rows must be visible
            o2, l2 = vo.ivf_lists(np.array(all_assign, np.uint64), C)
            ids, d, cnt = idx.search_batch(extra[:8], k, nprobe=8)
            oi, od, oc = vo.ivf_search(all_rows[: n + i], cents, o2, l2, extra[:8], k, nprobe=8)
            assert np.array_equal(ids, oi) and np.array_equal(bits(d), bits(od))
            assert all(ids[j][0] == n + j for j in range(8))  # each added row finds itself
        _, cl = idx.add(extra[i], 0)
        all_assign.append(cl)
    batch = data(vo, 500, dim, seed=6, n_centers=300)
    _, cl = idx.add_batch(batch)
    all_rows = np.vstack([all_rows, batch])
    all_assign = np.array(all_assign + [int(c) for c in cl], np.uint64)
    off, lr = vo.ivf_lists(all_assign, C)
    q2 = np.vstack([q[:60], extra[-4:], batch[:20]])
    ids, d, cnt = idx.search_batch(q2, k, nprobe=8)
    oi, od, oc = vo.ivf_search(all_rows, cents, off, lr, q2, k, nprobe=8)
    assert np.array_equal(cnt, oc) and np.array_equal(ids, oi) and np.array_equal(bits(d), bits(od))


def test_ivf_h16_candidate_copy_falls_back_on_ties(vb, vo, ctx):
    """duplicates beyond the candidate list and unnormalised rows of very different norms: the fp16 pass must flag
    what it cannot prove and the exact redo must return the reference's order"""
    n, dim, C, k = 6000, 96, 8, 10
    rows = data(vo, n, dim, normalize=False)
    rows[::7] *= np.float32(37.0)
    rows[1000:1100] = rows[17]
    init = vo.init_rows(3, 1, C, n)
    idx = vb.IVFFlatIndex.build_index(C, 1, 6, rows, init_rows=init, ctx=ctx)
    cents, assign, _, _ = vo.ivf_build_index(rows, C, 1, 6, init)
    off, lr = vo.ivf_lists(assign, C)
    q = data(vo, 40, dim, seed=2, normalize=False)
    q[0] = rows[17]
    idx.set_mode(4)
    ids, d, cnt = idx.search_batch(q, k, nprobe=4)
    st = idx.last_search_stats()
    oi, od, oc = vo.ivf_search(rows, cents, off, lr, q, k, nprobe=4)
    assert np.array_equal(ids, oi) and np.array_equal(bits(d), bits(od)) and np.array_equal(cnt, oc)
    assert list(ids[0]) == [17] + list(range(1000, 1009))
    assert st["uncertified_queries"] >= 1


def test_ivf_host_call_graph_replay_stays_exact(vb, vo, ctx):
    """vers_ivf_search replays its device work from a CUDA graph once a call shape repeats (call 1 eager, call 2
    captures, calls 3+ replay): every call must return the oracle's bits for ITS queries, a changed shape or a changed
    index (add) must start over, and both candidate modes must behave the same"""
    n, dim, C, k = 12000, 96, 32, 10
    rows = data(vo, n, dim, n_centers=200)
    init = vo.init_rows(3, 1, C, n)
    idx = vb.IVFFlatIndex.build_index(C, 1, 4, rows, init_rows=init, ctx=ctx)
    cents, assign, _, _ = vo.ivf_build_index(rows, C, 1, 4, init)
    off, lr = vo.ivf_lists(assign, C)
    for mode in (4, 0):
        idx.set_mode(mode)
        for rep in range(6):  # different queries behind the same shape: a replay must read the new batch
            q = data(vo, 64, dim, seed=10 + rep, n_centers=200)
            ids, d, cnt = idx.search_batch(q, k, nprobe=8)
            oi, od, oc = vo.ivf_search(rows, cents, off, lr, q, k, nprobe=8)
            assert np.array_equal(ids, oi) and np.array_equal(bits(d), bits(od)) and np.array_equal(cnt, oc), (mode, rep)
        q = data(vo, 33, dim, seed=30, n_centers=200)  # another shape in between, then back
        ids, d, cnt = idx.search_batch(q, 5, nprobe=4)
        oi, od, oc = vo.ivf_search(rows, cents, off, lr, q, 5, nprobe=4)
        assert np.array_equal(ids, oi) and np.array_equal(bits(d), bits(od))
    # the index changes under a warm graph: the added row must be found by the very next call of the same shape
    q = data(vo, 64, dim, seed=11, n_centers=200)
    for rep in range(3):
        idx.search_batch(q, k, nprobe=8)
    extra = (q[5] + np.float32(1e-3)).astype(np.float32)
    new_id, cl = idx.add(extra, 0)
    all_rows = np.vstack([rows, extra[None]])
    all_assign = np.concatenate([assign, np.array([cl], np.uint64)])
    off2, lr2 = vo.ivf_lists(all_assign, C)
    for rep in range(3):
        ids, d, cnt = idx.search_batch(q, k, nprobe=8)
        oi, od, oc = vo.ivf_search(all_rows, cents, off2, lr2, q, k, nprobe=8)
        assert np.array_equal(ids, oi) and np.array_equal(bits(d), bits(od))
    assert new_id in ids[5]


def test_ivf_search_approximate_single_query_trait_call(vo, ivf_c1):
    s = ivf_c1
    q = data(vo, 3, 300, seed=2)
    off, lr = vo.ivf_lists(s["assign"], s["C"])
    oi, od, oc = vo.ivf_search(s["rows"], s["cents"], off, lr, q, 10, nprobe=0)
    for i in range(3):
        got = s["idx"].search_approximate(q[i], 10)
        assert [g[0] for g in got] == list(oi[i][: oc[i]])
        assert np.array_equal(bits(np.array([g[1] for g in got], np.float32)), bits(od[i][: oc[i]]))


def test_ivf_spill_semantics_small_lists(vb, vo, ctx):
    """lists shorter than top_k force the reference's spill path (ivfflat.rs:181-185): concatenated, not re-sorted"""
    n, dim, C, k = 300, 32, 64, 20
    rows = data(vo, n, dim, n_centers=40)
    init = vo.init_rows(3, 1, C, n)
    idx = vb.IVFFlatIndex.build_index(C, 1, 5, rows, init_rows=init, ctx=ctx)
    cents, assign, _, _ = vo.ivf_build_index(rows, C, 1, 5, init)
    assert np.array_equal(idx.assignments, assign)
    off, lr = vo.ivf_lists(assign, C)
    q = data(vo, 64, dim, seed=2, n_centers=40)
    ids, d, cnt = idx.search_batch(q, k, nprobe=0)
    oi, od, oc = vo.ivf_search(rows, cents, off, lr, q, k, nprobe=0)
    assert np.array_equal(cnt, oc) and np.array_equal(ids, oi) and np.array_equal(bits(d), bits(od))
    # at least one query must actually have spilled, else the test is vacuous
    first = vo.assign(q, cents)
    sizes = np.bincount(assign.astype(np.int64), minlength=C)
    assert np.any(sizes[first.astype(np.int64)] < k)


def test_ivf_search_panics_like_reference_when_index_too_small(vb, vo, ctx):
    rows = data(vo, 8, 16)
    init = vo.init_rows(3, 1, 4, 8)
    idx = vb.IVFFlatIndex.build_index(4, 1, 3, rows, init_rows=init, ctx=ctx)
    with pytest.raises(vb.VersPanic):
        idx.search_approximate(rows[0], 9)  # index out of bounds at ivfflat.rs:169
    assert len(idx.search_approximate(rows[0], 8)) == 8


def test_ivf_add_then_search(vb, vo, ctx):
    n, dim, C = 2000, 64, 8
    rows = data(vo, n, dim)
    extra = data(vo, 300, dim, seed=5)
    init = vo.init_rows(3, 1, C, n)
    idx = vb.IVFFlatIndex.build_index(C, 1, 10, rows, init_rows=init, ctx=ctx)
    cents, assign, _, _ = vo.ivf_build_index(rows, C, 1, 10, init)
    all_rows = np.vstack([rows, extra])
    all_assign = list(assign)
    for i in range(extra.shape[0]):
        new_id, cl = idx.add(extra[i], vec_id=999999)  # the caller's id is ignored (ivfflat.rs:209)
        assert new_id == n + i
        assert cl == vo.nearest_centroid(cents, extra[i])
        all_assign.append(cl)
    all_assign = np.array(all_assign, np.uint64)
    assert len(idx) == n + 300
    assert np.array_equal(idx.assignments, all_assign)
    bad = extra[0].copy()
    bad[3] = np.nan  # partial_cmp(..).unwrap() panics on the NaN distance (ivfflat.rs:207): nothing may change
    with pytest.raises(vb.VersPanic):
        idx.add(bad, 0)
    assert len(idx) == n + 300 and np.array_equal(idx.assignments, all_assign)
    off, lr = vo.ivf_lists(all_assign, C)
    q = data(vo, 50, dim, seed=2)
    for nprobe in (0, 3):
        ids, d, cnt = idx.search_batch(q, 10, nprobe=nprobe)
        oi, od, oc = vo.ivf_search(all_rows, cents, off, lr, q, 10, nprobe=nprobe)
        assert np.array_equal(ids, oi) and np.array_equal(bits(d), bits(od)) and np.array_equal(cnt, oc)


def test_ivf_add_batch_equals_sequential_adds(vb, vo, ctx):
    """vers_ivf_add_batch == the same embeddings added one by one (ivfflat.rs:200-213): ids, clusters, list contents and
    every later search; large enough to overflow the lists' slack (one re-layout); a NaN row panics and changes nothing"""
    n, dim, C = 3000, 64, 12
    rows = data(vo, n, dim)
    extra = data(vo, 700, dim, seed=5)
    init = vo.init_rows(3, 1, C, n)
    a = vb.IVFFlatIndex.build_index(C, 1, 6, rows, init_rows=init, ctx=ctx)
    b = vb.IVFFlatIndex.build_index(C, 1, 6, rows, init_rows=init, ctx=ctx)
    ids, cl = a.add_batch(extra)
    seq = [b.add(extra[i], 12345) for i in range(extra.shape[0])]
    assert np.array_equal(ids, np.arange(n, n + 700, dtype=np.uint64)) and [int(x) for x in ids] == [s[0] for s in seq]
    assert [int(x) for x in cl] == [s[1] for s in seq]
    cents = a.centroids
    assert np.array_equal(cl.astype(np.uint64), vo.assign(extra, cents))  # nearest centroid, first minimum
    assert np.array_equal(a.assignments, b.assignments) and np.array_equal(a.list_sizes, b.list_sizes)
    for c in range(C):
        ia, ra = a.get_list(c, with_rows=True)
        ib, rb = b.get_list(c, with_rows=True)
        assert np.array_equal(ia, ib) and np.array_equal(bits(ra), bits(rb))
    q = np.vstack([extra[:40], data(vo, 40, dim, seed=2)])
    allrows = np.vstack([rows, extra])
    off, lr = vo.ivf_lists(a.assignments, C)
    for nprobe in (0, 3):
        ga = a.search_batch(q, 5, nprobe=nprobe)
        gb = b.search_batch(q, 5, nprobe=nprobe)
        oi, od, oc = vo.ivf_search(allrows, cents, off, lr, q, 5, nprobe=nprobe)
        assert np.array_equal(ga[0], gb[0]) and np.array_equal(bits(ga[1]), bits(gb[1]))
        assert np.array_equal(ga[0], oi) and np.array_equal(bits(ga[1]), bits(od)) and np.array_equal(ga[2], oc)
    bad = extra[:5].copy()
    bad[2, 7] = np.nan
    sizes = a.list_sizes.copy()
    with pytest.raises(vb.VersPanic):
        a.add_batch(bad)
    assert np.array_equal(a.list_sizes, sizes) and len(a) == n + 700


def test_ivf_from_parts_and_save_load_roundtrip(vb, vo, ctx, tmp_path, ivf_c1):
    s = ivf_c1
    p = str(tmp_path / "ivf.bin")
    s["idx"].save_index(p)
    idx2 = vb.IVFFlatIndex.load_index(p, ctx=ctx)
    q = data(vo, 20, 300, seed=2)
    a = s["idx"].search_batch(q, 10, nprobe=0)
    b = idx2.search_batch(q, 10, nprobe=0)
    assert np.array_equal(a[0], b[0]) and np.array_equal(bits(a[1]), bits(b[1]))
    assert np.array_equal(idx2.assignments, s["assign"])


def test_ivf_recall_identical_to_oracle(vo, ivf_c1):
    s = ivf_c1
    q = data(vo, 100, 300, seed=2)
    off, lr = vo.ivf_lists(s["assign"], s["C"])
    gt, _, _ = vo.exhaustive(s["rows"], q, 10)
    ids, _, _ = s["idx"].search_batch(q, 10, nprobe=4)
    oi, _, _ = vo.ivf_search(s["rows"], s["cents"], off, lr, q, 10, nprobe=4)
    rec = lambda a: np.mean([len(set(a[i]) & set(gt[i])) / 10 for i in range(len(gt))])
    assert rec(ids) == rec(oi)


# ---------------------------------------------------------------------------------------------- merge
def test_topk_merge_by_distance_then_id(vb, ctx):
    import ctypes as C

    import torch

    parts, nq, k = 3, 50, 10
    rng = np.random.default_rng(0)
    d = rng.integers(0, 20, (parts, nq, k)).astype(np.float32)  # many ties
    d.sort(axis=2)
    ids = rng.permutation(parts * nq * k).reshape(parts, nq, k).astype(np.uint64)
    ids[2, :, 7:] = np.iinfo(np.uint64).max  # short lists
    d[2, :, 7:] = np.inf
    t_ids = torch.from_numpy(ids.view(np.int64)).cuda()
    t_d = torch.from_numpy(d).cuda()
    o_ids = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    o_d = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    o_c = torch.empty((nq,), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    vb._abi.check(vb.lib().vers_topk_merge_dev(ctx.h, C.c_void_p(t_ids.data_ptr()), C.c_void_p(t_d.data_ptr()), parts,
                                               0, 0, nq, k, C.c_void_p(o_ids.data_ptr()), C.c_void_p(o_d.data_ptr()),
                                               C.c_void_p(o_c.data_ptr())))
    ctx.sync()
    got = o_ids.cpu().numpy().view(np.uint64)
    for q in range(nq):
        cand = [(d[p, q, e], ids[p, q, e]) for p in range(parts) for e in range(k) if ids[p, q, e] != np.iinfo(np.uint64).max]
        cand.sort()
        assert [c[1] for c in cand[:k]] == list(got[q])


# ---------------------------------------------------------------------------------------------- LSH hashing
@pytest.mark.parametrize("n,dim,P", [(10000, 300, 16), (3000, 128, 100), (500, 33, 1)])
def test_lsh_hash_bits_exact(vb, vo, ctx, n, dim, P):
    rows = data(vo, n, dim)
    rng = np.random.default_rng(1)
    planes = np.empty((P, dim), np.float32)
    consts = np.empty(P, np.float32)
    for p in range(P):
        a, b = rng.choice(n, 2, replace=False)
        planes[p], consts[p] = vo.lsh_make_plane(rows[a], rows[b])
    ds = vb.Dataset.upload(ctx, rows)
    got = vb.lsh_hash(ds, planes, consts)
    assert np.array_equal(got, vo.lsh_hash(rows, planes, consts))
    assert 0.2 < got.mean() < 0.8


def test_no_cpu_fallback_symbols(vb):
    # the product library must not depend on the oracle
    import subprocess

    out = subprocess.run(["ldd", vb.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


@pytest.mark.parametrize("gpus", [2, 4, 8])
def test_multi_gpu_sharded_parity(gpus):
    """runs tests/mgpu_check.py under torchrun on `gpus` GPUs (skipped when the box has fewer): chained k-means,
    list- and row-sharded search through vers_comm / vers_sharded_* against the single-process oracle"""
    import os
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(gpus),
                        "--master-addr", "127.0.0.1", "--master-port", str(29517 + gpus),
                        os.path.join(root, "tests", "mgpu_check.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "mgpu_check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_comm_world1_is_the_single_gpu_path(vb, vo, ctx):
    """vers_comm with world == 1 (no NCCL needed): the vers_sharded_* calls are the plain single-GPU calls"""
    from vers_b200.sharded import Comm, ShardedIVFFlat

    n, dim, C, k = 6000, 128, 24, 10
    rows = data(vo, n, dim)
    q = data(vo, 40, dim, seed=2)
    init = vo.init_rows(3, 1, C, n)[0]
    comm = Comm(ctx, 0, 1)
    ds = vb.Dataset.upload(ctx, rows)
    index = ShardedIVFFlat.build(comm, ds, C, 6, init)
    cents, assign, _ = vo.kmeans_fit(rows, init, 6)
    assert np.array_equal(index.ivf.assignments, assign) and np.array_equal(bits(index.ivf.centroids), bits(cents))
    off, lr = vo.ivf_lists(assign, C)
    ids, d, cnt = index.search(q, k, 4)
    oi, od, oc = vo.ivf_search(rows, cents, off, lr, q, k, nprobe=4)
    assert np.array_equal(ids, oi) and np.array_equal(bits(d), bits(od)) and np.array_equal(cnt, oc)
    comm.close()


# ---------------------------------------------------------------------------------------------- LSH forest
def _same_forest(vb_idx, vo_idx, T):
    for t in range(T):
        a, b = vb_idx.flatten(t), vo_idx.flatten(t)
        assert np.array_equal(a["kind"], b["kind"]), f"tree {t}: shape differs"
        assert np.array_equal(a["leaf_len"], b["leaf_len"])
        assert np.array_equal(bits(a["planes"]), bits(b["planes"])), f"tree {t}: plane coefficient bits differ"
        assert np.array_equal(bits(a["consts"]), bits(b["consts"]))
        assert np.array_equal(a["items"], b["items"]), f"tree {t}: leaf members / order differ"


@pytest.mark.parametrize("n,dim,T,max_size", [(6000, 300, 4, 100), (3000, 24, 6, 12), (500, 33, 3, 2)])
def test_lsh_forest_build_identical_to_oracle(vb, vo, ctx, n, dim, T, max_size):
    rows = data(vo, n, dim, n_centers=20)
    rows[100] = rows[7]  # a duplicate: dropped by deduplicate (lsh.rs:113-130)
    ids = np.arange(n, dtype=np.uint64) + 1000
    g = vb.ANNIndex.build_index(T, max_size, rows, ids, seed=4, ctx=ctx)
    o = vo.LSH(rows, ids, T, max_size, 4)
    assert g.info()["num_values"] == o.num_values == n - 1
    _same_forest(g, o, T)


@pytest.mark.parametrize("k", [1, 10, 50])
def test_lsh_search_identical_to_oracle(vb, vo, ctx, k):
    n, dim, T, max_size = 8000, 300, 8, 100
    rows = data(vo, n, dim, n_centers=30)
    g = vb.ANNIndex.build_index(T, max_size, rows, None, seed=4, ctx=ctx)
    o = vo.LSH(rows, None, T, max_size, 4)
    q = data(vo, 120, dim, seed=2, n_centers=30)
    ids, d, cnt = g.search_batch(q, k)
    oi, od, oc = o.search(q, k)
    assert np.array_equal(cnt, oc)
    assert np.array_equal(ids, oi)
    assert np.array_equal(bits(d), bits(od))
    got = g.search_approximate(q[0], k)
    assert [x[0] for x in got] == list(oi[0][: oc[0]])


def test_lsh_search_backtracks_through_small_leaves(vb, vo, ctx):
    """max_size smaller than top_k: every leaf is short, tree_result backtracks all the way up (lsh.rs:203-213)"""
    n, dim, T, max_size, k = 1500, 16, 5, 6, 20
    rows = data(vo, n, dim, n_centers=10)
    g = vb.ANNIndex.build_index(T, max_size, rows, None, seed=9, ctx=ctx)
    o = vo.LSH(rows, None, T, max_size, 9)
    _same_forest(g, o, T)
    q = data(vo, 64, dim, seed=2, n_centers=10)
    ids, d, cnt = g.search_batch(q, k)
    oi, od, oc = o.search(q, k)
    assert np.array_equal(cnt, oc) and np.array_equal(ids, oi) and np.array_equal(bits(d), bits(od))


def test_lsh_add_splits_and_search(vb, vo, ctx):
    n, dim, T, max_size = 600, 32, 3, 10
    rows = data(vo, n, dim, n_centers=8)
    extra = data(vo, 250, dim, seed=5, n_centers=8)
    g = vb.ANNIndex.build_index(T, max_size, rows, None, seed=4, ctx=ctx)
    o = vo.LSH(rows, None, T, max_size, 4)
    for i in range(extra.shape[0]):
        g.add(extra[i], n + i)
        o.add(extra[i], n + i)
    assert g.info()["num_values"] == o.num_values == n + 250
    _same_forest(g, o, T)
    q = np.vstack([extra[:20], data(vo, 20, dim, seed=2, n_centers=8)])
    ids, d, cnt = g.search_batch(q, 5)
    oi, od, oc = o.search(q, 5)
    assert np.array_equal(ids, oi) and np.array_equal(bits(d), bits(od)) and np.array_equal(cnt, oc)
    assert [int(ids[i][0]) for i in range(20)] == [n + i for i in range(20)]
    with pytest.raises(vb.VersPanic):
        g.add(extra[0], 10**6)  # vec_id is stored as a row index (lsh.rs:247): out of range


def test_lsh_1m_shape_hash_parity_sample(vb, vo, ctx):
    """BASELINE config 3 shape (scaled rows): 16 planes over many rows, every bit equal"""
    n, dim, P = 200000, 300, 16
    rows = data(vo, n, dim, n_centers=64)
    rng = np.random.default_rng(2)
    planes = np.empty((P, dim), np.float32)
    consts = np.empty(P, np.float32)
    for p in range(P):
        a, b = rng.choice(n, 2, replace=False)
        planes[p], consts[p] = vo.lsh_make_plane(rows[a], rows[b])
    ds = vb.Dataset.upload(ctx, rows)
    assert np.array_equal(vb.lsh_hash(ds, planes, consts), vo.lsh_hash(rows, planes, consts))


def test_cpp_host_mirror(vb, tmp_path):
    """host/vers_index.hpp (C++ mirror of Index<N> / IVFFlatIndex / ANNIndex over the C ABI) end to end"""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "host_check")
    libdir = os.path.dirname(vb.LIB_PATH)
    subprocess.run(["g++", "-std=c++17", "-O1", os.path.join(root, "host", "host_check.cpp"), "-o", exe, "-L" + libdir,
                    "-lvers_b200", "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([exe, str(tmp_path / "ivf.bin")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "host_check ok" in r.stdout, r.stdout + r.stderr


def test_lsh_save_index_writes_the_reference_layout(vb, vo, ctx, tmp_path):
    """ANNIndex.save_index: the device forest serialized in the reference's bincode layout (lsh.rs:13-55) equals the
    oracle's forest serialized the same way, byte for byte"""
    from vers_b200.bincode import read_ann, write_ann

    n, dim, T, max_size = 3000, 24, 4, 12
    rows = data(vo, n, dim, n_centers=20)
    rows[100] = rows[7]
    ids = np.arange(n, dtype=np.uint64) + 1000
    g = vb.ANNIndex.build_index(T, max_size, rows, ids, seed=4, ctx=ctx)
    o = vo.LSH(rows, ids, T, max_size, 4)
    pg, po = str(tmp_path / "g.bin"), str(tmp_path / "o.bin")
    g.save_index(pg)
    keep = np.array([i for i in range(n) if i != 100])
    write_ann(po, max_size, [o.flatten(t) for t in range(T)], rows[keep], ids[keep])
    assert open(pg, "rb").read() == open(po, "rb").read()
    mns, trees, values, rid = read_ann(pg, dim)
    assert mns == max_size and len(trees) == T and values.shape == (n - 1, dim) and np.array_equal(rid, ids[keep])


def test_lsh_load_index_roundtrip_after_adds(vb, vo, ctx, tmp_path):
    """the reference's run_test order: build_index, add, save_index, load_index (base.rs:31-58): the loaded forest is
    the same forest (structure, plane bits, leaf members), searches identically, still equals the oracle that went
    through the same adds, and keeps accepting adds (leaf splits draw their sample pairs from the node path hash)"""
    n, dim, T, max_size = 900, 40, 3, 10
    rows = data(vo, n, dim, n_centers=8)
    extra = data(vo, 120, dim, seed=5, n_centers=8)
    g = vb.ANNIndex.build_index(T, max_size, rows, None, seed=4, ctx=ctx)
    o = vo.LSH(rows, None, T, max_size, 4)
    for i in range(60):
        g.add(extra[i], n + i)
        o.add(extra[i], n + i)
    path = str(tmp_path / "ann.bin")
    g.save_index(path)
    h = vb.ANNIndex.load_index(path, dim, seed=4, ctx=ctx)
    assert h.info() == g.info() and h.max_node_size == max_size
    _same_forest(h, o, T)
    v, ids = h.values_and_ids()
    assert np.array_equal(bits(v), bits(np.vstack([rows, extra[:60]]))) and np.array_equal(ids, np.arange(n + 60))
    q = np.vstack([extra[:10], data(vo, 30, dim, seed=2, n_centers=8)])
    hi, hd, hc = h.search_batch(q, 7)
    gi, gd, gc = g.search_batch(q, 7)
    oi, od, oc = o.search(q, 7)
    assert np.array_equal(hi, gi) and np.array_equal(bits(hd), bits(gd)) and np.array_equal(hc, gc)
    assert np.array_equal(hi, oi) and np.array_equal(bits(hd), bits(od)) and np.array_equal(hc, oc)
    for i in range(60, 120):  # the loaded index and the original keep evolving identically
        g.add(extra[i], n + i)
        h.add(extra[i], n + i)
        o.add(extra[i], n + i)
    _same_forest(h, o, T)
    _same_forest(g, o, T)
    # a rejected add leaves the index untouched (vec_id is used as a row index, lsh.rs:247)
    before = h.info()
    with pytest.raises(vb.VersPanic):
        h.add(extra[0], 10**6)
    assert h.info() == before
    path2 = str(tmp_path / "ann2.bin")
    h.save_index(path2)
    g.save_index(path)
    assert open(path, "rb").read() == open(path2, "rb").read()
