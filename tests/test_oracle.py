"""Pins the CPU oracle (oracle/vers_oracle.c).

The reference has no tests, fixtures or golden vectors, so the oracle is pinned three ways:
  1. an INDEPENDENT numpy restatement of the reference arithmetic (column-at-a-time float32 accumulation, which is
     exactly the reference's left-to-right order with rounded sub/mul/add) must agree with the C oracle bit for bit;
  2. float64 sanity bounds (catches a wrong formula that happens to be self-consistent);
  3. the committed golden vectors (tests/golden/, produced by tests/golden/make_golden.py) must reproduce.
"""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


# ---- independent numpy restatement -------------------------------------------------------------------------
def np_l2sq(rows, q):
    """base.rs:119-126: s = 0; s = s + (a-b)*(a-b), left to right, float32"""
    acc = np.zeros(rows.shape[0], np.float32)
    for j in range(rows.shape[1]):
        t = (rows[:, j] - q[j]).astype(np.float32)
        acc = (acc + (t * t).astype(np.float32)).astype(np.float32)
    return acc


def np_dot(rows, q):
    """base.rs:91-93"""
    acc = np.zeros(rows.shape[0], np.float32)
    for j in range(rows.shape[1]):
        acc = (acc + (rows[:, j] * q[j]).astype(np.float32)).astype(np.float32)
    return acc


def np_normalize(rows):
    """base.rs:95-105"""
    m = np.zeros(rows.shape[0], np.float32)
    for j in range(rows.shape[1]):
        m = (m + (rows[:, j] * rows[:, j]).astype(np.float32)).astype(np.float32)
    m = np.sqrt(m).astype(np.float32)
    out = rows.copy()
    ok = ~(m < np.float32(1e-6))
    out[ok] = (rows[ok] / m[ok, None]).astype(np.float32)
    return out


def np_assign(rows, cents):
    """ivfflat.rs:29-46: first minimum"""
    d = np.stack([np_l2sq(rows, c) for c in cents], axis=1)
    return np.argmin(d, axis=1).astype(np.uint64)  # numpy argmin returns the first minimum


def np_update(rows, assign, C):
    """ivfflat.rs:47-71: row-order sums from +0.0, / count as f32, zero vector when empty"""
    sums = np.zeros((C, rows.shape[1]), np.float32)
    counts = np.zeros(C, np.uint64)
    for r in range(rows.shape[0]):
        c = int(assign[r])
        sums[c] = (sums[c] + rows[r]).astype(np.float32)
        counts[c] += 1
    out = np.zeros_like(sums)
    for c in range(C):
        if counts[c] > 0:
            out[c] = (sums[c] / np.float32(counts[c])).astype(np.float32)
    return out, counts


def np_kmeans(rows, init, max_iter):
    """ivfflat.rs:73-100"""
    cents = rows[init.astype(np.int64)].copy()
    iters = 0
    for _ in range(max_iter):
        a = np_assign(rows, cents)
        new, _ = np_update(rows, a, cents.shape[0])
        iters += 1
        if np.array_equal(bits(new), bits(cents)):
            break
        cents = new
    return cents, np_assign(rows, cents), iters


def stable_topk(d, ids, k):
    order = np.argsort(d, kind="stable")[:k]
    return ids[order], d[order]


def np_ivf_search_ref(rows, cents, lists, q, k):
    """ivfflat.rs:153-198"""
    cd = np_l2sq(cents, q)
    nearest = np.argsort(cd, kind="stable")
    out_i, out_d = [], []
    curr, remainder = 0, k
    while len(out_i) < k:
        if curr >= len(nearest):
            raise IndexError("panic: index out of bounds")
        lst = lists[nearest[curr]]
        d = np_l2sq(rows[lst.astype(np.int64)], q) if len(lst) else np.zeros(0, np.float32)
        pi, pd = stable_topk(d, lst, k)
        if len(pi) < remainder:
            remainder -= len(pi)
            out_i += list(pi)
            out_d += list(pd)
            curr += 1
        else:
            out_i += list(pi[:remainder])
            out_d += list(pd[:remainder])
            break
    return np.array(out_i, np.uint64), np.array(out_d, np.float32)


# ---- tests ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim", [1, 7, 128, 300, 768])
def test_l2sq_dot_cosine_bits_vs_numpy(vo, dim):
    rows = vo.synth(1, 64, dim, kind=1, n_centers=4, center_seed=7)
    q = vo.synth(2, 1, dim, kind=1, n_centers=4, center_seed=7)[0]
    l2 = np.array([vo.l2sq(r, q) for r in rows], np.float32)
    dt = np.array([vo.dot(r, q) for r in rows], np.float32)
    cs = np.array([vo.cosine_distance(r, q) for r in rows], np.float32)
    assert np.array_equal(bits(l2), bits(np_l2sq(rows, q)))
    assert np.array_equal(bits(dt), bits(np_dot(rows, q)))
    assert np.array_equal(bits(cs), bits((np.float32(1.0) - np_dot(rows, q)).astype(np.float32)))
    # float64 sanity: within the worst-case fp32 summation bound
    ref = ((rows.astype(np.float64) - q.astype(np.float64)) ** 2).sum(1)
    assert np.all(np.abs(l2 - ref) <= (dim + 3) * 2.0 ** -23 * np.maximum(ref, 1e-30))


def test_l2sq_is_not_the_fma_or_pairwise_value(vo):
    """guards the build flags: a contracted or re-associated build gives different low bits on this input"""
    rows = vo.synth(1, 2000, 300)
    q = vo.synth(2, 1, 300)[0]
    seq = np.array([vo.l2sq(r, q) for r in rows], np.float32)
    pairwise = ((rows - q) ** 2).sum(1, dtype=np.float32)  # numpy pairwise summation
    assert np.array_equal(bits(seq), bits(np_l2sq(rows, q)))
    assert np.mean(bits(seq) != bits(pairwise)) > 0.3


@pytest.mark.parametrize("kind", [0, 1])
def test_synth_matches_spec_and_normalize(vo, kind):
    n, dim = 50, 37
    raw = vo.synth(5, n, dim, kind=kind, n_centers=3, center_seed=8, row0=10, normalize=False)
    M = (1 << 64) - 1

    def sm(x):
        x = (x + 0x9E3779B97F4A7C15) & M
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M
        return x ^ (x >> 31)

    def unit(seed, i):
        return np.float32(((sm((sm(seed) + i) & M) >> 40) - 8388608) / 8388608.0)

    for r in range(0, n, 7):
        for c in range(0, dim, 5):
            u = unit(5, (10 + r) * dim + c)
            if kind == 1:
                ctr = sm(8 ^ 0x5BD1E995 ^ (((10 + r) * 0x9E3779B97F4A7C15) & M)) % 3
                u = np.float32(unit(8, ctr * dim + c) + np.float32(0.25) * u)
            assert raw[r, c] == u
    assert np.all(np.abs(raw) <= 1.25)
    nz = vo.synth(5, n, dim, kind=kind, n_centers=3, center_seed=8, row0=10, normalize=True)
    assert np.array_equal(bits(nz), bits(np_normalize(raw)))
    assert np.array_equal(bits(vo.normalize_rows(raw)), bits(np_normalize(raw)))


def test_normalize_leaves_tiny_rows(vo):
    rows = np.zeros((3, 8), np.float32)
    rows[1] = 1e-8
    rows[2] = 3.0
    out = vo.normalize_rows(rows)
    assert np.array_equal(out[0], rows[0]) and np.array_equal(out[1], rows[1])
    assert abs(np.linalg.norm(out[2]) - 1) < 1e-6


def test_assign_update_kmeans_vs_numpy(vo):
    n, dim, C = 600, 40, 9
    rows = vo.synth(1, n, dim, kind=1, n_centers=6, center_seed=7)
    init = vo.init_rows(3, 1, C, n)[0]
    cents0 = rows[init.astype(np.int64)]
    a = vo.assign(rows, cents0)
    assert np.array_equal(a, np_assign(rows, cents0))
    c1, cnt = vo.update(rows, a, C)
    n1, ncnt = np_update(rows, a, C)
    assert np.array_equal(cnt, ncnt) and np.array_equal(bits(c1), bits(n1))
    cents, assign, iters = vo.kmeans_fit(rows, init, 12)
    nc, na, ni = np_kmeans(rows, init, 12)
    assert iters == ni and np.array_equal(assign, na) and np.array_equal(bits(cents), bits(nc))
    cost = vo.kmeans_cost(rows, cents, assign)
    acc = np.float32(0)
    for r in range(n):
        acc = np.float32(acc + np_l2sq(rows[r:r + 1], cents[int(assign[r])])[0])
    assert bits(cost) == bits(acc)


def test_duplicate_init_rows_give_empty_cluster_zero_centroid(vo):
    """draws WITH replacement (ivfflat.rs:23) -> duplicate centroids -> the later twin never wins (first minimum)
    -> empty cluster -> zero vector (ivfflat.rs:63-67)"""
    rows = vo.synth(1, 200, 16)
    init = np.array([5, 9, 5], np.uint64)
    a0 = vo.assign(rows, rows[init.astype(np.int64)])
    assert not np.any(a0 == 2)
    c1, cnt = vo.update(rows, a0, 3)
    assert cnt[2] == 0 and not c1[2].any()
    # build_kmeans adopts the zero vector as centroid 2 and carries on (ivfflat.rs:95)
    cents, assign, _ = vo.kmeans_fit(rows, init, 1)
    assert np.array_equal(bits(cents), bits(c1)) and np.array_equal(assign, vo.assign(rows, c1))


def test_update_sharded_order_differs_only_in_association(vo):
    rows = vo.synth(1, 5000, 32, normalize=False)
    a = (np.arange(5000) % 7).astype(np.uint64)
    c1, n1 = vo.update(rows, a, 7)
    c4, n4 = vo.update(rows, a, 7, shards=4)
    assert np.array_equal(n1, n4)
    assert np.allclose(c1, c4, rtol=1e-5, atol=1e-6)
    c1b, _ = vo.update(rows, a, 7, shards=1)
    assert np.array_equal(bits(c1), bits(c1b))


def test_build_index_best_of_attempts_strict_less(vo):
    rows = vo.synth(1, 400, 24, kind=1, n_centers=5, center_seed=7)
    init = vo.init_rows(3, 3, 6, 400)
    cents, assign, cost, best = vo.ivf_build_index(rows, 6, 3, 8, init)
    costs = []
    for t in range(3):
        c, a, _ = vo.kmeans_fit(rows, init[t], 8)
        costs.append(vo.kmeans_cost(rows, c, a))
    assert best == int(np.argmin(costs))  # first of equal minima == strict `<`
    assert bits(cost) == bits(np.float32(min(costs)))
    # identical attempts: the FIRST one must be kept (strict <, ivfflat.rs:116)
    init2 = np.stack([init[0], init[0]])
    assert vo.ivf_build_index(rows, 6, 2, 8, init2)[3] == 0


def test_ivf_search_reference_semantics_vs_numpy(vo):
    n, dim, C, k = 500, 20, 40, 15
    rows = vo.synth(1, n, dim, kind=1, n_centers=25, center_seed=7)
    init = vo.init_rows(3, 1, C, n)
    cents, assign, _, _ = vo.ivf_build_index(rows, C, 1, 6, init)
    off, lr = vo.ivf_lists(assign, C)
    lists = [lr[off[c]:off[c + 1]] for c in range(C)]
    for c in range(C):  # ids[c] ascending (ivfflat.rs:123-127)
        assert np.all(np.diff(lists[c].astype(np.int64)) > 0)
    q = vo.synth(2, 30, dim, kind=1, n_centers=25, center_seed=7)
    ids, d, cnt = vo.ivf_search(rows, cents, off, lr, q, k, nprobe=0)
    spilled = 0
    for i in range(q.shape[0]):
        ni, nd = np_ivf_search_ref(rows, cents, lists, q[i], k)
        assert cnt[i] == len(ni) == k
        assert np.array_equal(ids[i], ni) and np.array_equal(bits(d[i]), bits(nd))
        spilled += int(np.any(np.diff(nd) < 0))  # concatenation is not globally sorted
    assert spilled > 0


def test_ivf_search_panics_when_fewer_rows_than_k(vo):
    rows = vo.synth(1, 8, 6)
    cents, assign, _, _ = vo.ivf_build_index(rows, 3, 1, 3, vo.init_rows(3, 1, 3, 8))
    off, lr = vo.ivf_lists(assign, 3)
    with pytest.raises(vo.OraclePanic):
        vo.ivf_search(rows, cents, off, lr, rows[:1], 9, nprobe=0)
    ids, d, cnt = vo.ivf_search(rows, cents, off, lr, rows[:1], 8, nprobe=0)
    assert cnt[0] == 8 and sorted(ids[0]) == list(range(8))


def test_ivf_nprobe_all_lists_equals_exhaustive(vo):
    rows = vo.synth(1, 700, 16, kind=1, n_centers=9, center_seed=7)
    cents, assign, _, _ = vo.ivf_build_index(rows, 12, 1, 5, vo.init_rows(3, 1, 12, 700))
    off, lr = vo.ivf_lists(assign, 12)
    q = vo.synth(2, 20, 16, kind=1, n_centers=9, center_seed=7)
    a = vo.ivf_search(rows, cents, off, lr, q, 10, nprobe=12)
    b = vo.exhaustive(rows, q, 10)
    assert np.array_equal(a[0], b[0]) and np.array_equal(bits(a[1]), bits(b[1]))


def test_exhaustive_stable_ties_and_metrics(vo):
    rows = vo.synth(1, 300, 12)
    rows[50:60] = rows[3]
    q = rows[3:4].copy()
    ids, d, cnt = vo.exhaustive(rows, q, 12)
    assert list(ids[0][:11]) == [3] + list(range(50, 60))
    ids2, d2, _ = vo.exhaustive(rows, vo.synth(2, 4, 12), 5, metric=1)
    for i in range(4):
        dd = (np.float32(1.0) - np_dot(rows, vo.synth(2, 4, 12)[i])).astype(np.float32)
        wi, wd = stable_topk(dd, np.arange(300, dtype=np.uint64), 5)
        assert np.array_equal(ids2[i], wi) and np.array_equal(bits(d2[i]), bits(wd))


def test_lsh_plane_and_hash_vs_numpy(vo):
    rows = vo.synth(1, 400, 30)
    coef, const = vo.lsh_make_plane(rows[3], rows[8])
    assert np.array_equal(bits(coef), bits((rows[8] - rows[3]).astype(np.float32)))  # other - self, base.rs:85-89
    mid = ((rows[3] + rows[8]).astype(np.float32) / np.float32(2.0)).astype(np.float32)
    assert bits(const) == bits(np.float32(-np_dot(coef[None, :], mid)[0]))
    planes = np.stack([coef, -coef])
    consts = np.array([const, -const], np.float32)
    got = vo.lsh_hash(rows, planes, consts)
    want = np.stack([(np_dot(rows, planes[p]) + consts[p]).astype(np.float32) >= 0 for p in range(2)], 1)
    assert np.array_equal(got.astype(bool), want)
    assert got[3, 0] == 0 and got[8, 0] == 1  # a is below, b is above its own mid-plane


def test_lsh_forest_structure_and_search_properties(vo):
    n, dim, T, max_size, k = 3000, 24, 5, 40, 10
    rows = vo.synth(1, n, dim, kind=1, n_centers=20, center_seed=7)
    rows[100] = rows[7]  # duplicate -> dropped by deduplicate (lsh.rs:113-130)
    ids = np.arange(n, dtype=np.uint64) + 1000
    L = vo.LSH(rows, ids, T, max_size, 4)
    assert L.num_values == n - 1
    for t in range(T):
        f = L.flatten(t)
        leaf = f["kind"] == 1
        assert np.all(f["leaf_len"][leaf] < max_size)  # lsh.rs:97
        assert sorted(f["items"]) == list(range(n - 1))  # every row in exactly one leaf
        assert f["planes"].shape[0] == (~leaf).sum() == leaf.sum() - 1
    q = vo.synth(2, 25, dim, kind=1, n_centers=20, center_seed=7)
    got_ids, got_d, cnt = L.search(q, k)
    dedup = np.delete(rows, 100, axis=0)
    dedup_ids = np.delete(ids, 100)
    for i in range(q.shape[0]):
        cand = L.candidates(q[i], k)
        assert len(cand) >= k
        d = np_l2sq(dedup[cand], q[i])
        wi, wd = stable_topk(d, cand.astype(np.uint64), k)
        assert np.array_equal(got_ids[i], dedup_ids[wi.astype(np.int64)])
        assert np.array_equal(bits(got_d[i]), bits(wd))
    # recall sanity against the exhaustive ground truth
    gt, _, _ = vo.exhaustive(dedup, q, k)
    rec = np.mean([len(set(got_ids[i] - 1000 - (got_ids[i] - 1000 > 100)) & set(gt[i])) / k for i in range(len(gt))])
    assert rec > 0.5


def test_lsh_add_splits_full_leaf(vo):
    n, dim = 300, 8
    rows = vo.synth(1, n, dim)
    L = vo.LSH(rows, None, 2, 10, 4)
    before = [len(L.flatten(t)["kind"]) for t in range(2)]
    extra = vo.synth(9, 200, dim)
    for i in range(200):
        L.add(extra[i], n + i)
    assert L.num_values == n + 200
    for t in range(2):
        f = L.flatten(t)
        assert len(f["kind"]) > before[t]
        assert sorted(f["items"]) == list(range(n + 200))
        assert np.all(f["leaf_len"][f["kind"] == 1] <= 10)  # a leaf may reach max_size after add (lsh.rs:240-248)
    ids, d, cnt = L.search(extra[:5], 3)
    assert [int(ids[i][0]) for i in range(5)] == [n + i for i in range(5)] and np.all(d[:, 0] == 0)


def np_simd_distance(u, v, metric):
    """independent restatement of base.rs:158-294: numpy float32 scalars, chunk sums by an explicit left-to-right loop
    (std::simd reduce_sum is the ordered reduction), chunk sums added in order"""
    dim = u.shape[0]
    p = ((u - v) * (u - v)).astype(np.float32) if metric == 0 else (u * v).astype(np.float32)
    res = np.float32(0.0)

    def chunk(lo, hi):
        s = np.float32(0.0)
        for i in range(lo, hi):
            s = np.float32(s + p[i])
        return s

    n64 = dim // 64
    for c in range(n64):
        res = np.float32(res + chunk(c * 64, c * 64 + 64))
    n4 = (dim - n64 * 64) // 4
    for c in range(n4):
        res = np.float32(res + chunk(n64 * 64 + c * 4, n64 * 64 + c * 4 + 4))
    for i in range(n64 * 64 + n4 * 4, dim):
        res = np.float32(res + p[i])
    return res if metric == 0 else np.float32(np.float32(1.0) - res)


@pytest.mark.parametrize("dim", [1, 3, 4, 63, 64, 67, 128, 300, 768])
def test_hnsw_simd_distances_vs_numpy(vo, dim):
    """the HNSW distance (SIMD association) differs from the scalar left-to-right one in the low bits and must match the
    numpy restatement exactly; panics on a missing row like id_to_vec.get(..).unwrap()"""
    rows = vo.synth(1, 40, dim, kind=1, n_centers=4, center_seed=7, normalize=True)
    q = vo.synth(2, 3, dim, kind=1, n_centers=4, center_seed=7, normalize=True)
    pr = np.arange(40, dtype=np.uint64)
    for metric in (0, 1):
        for qi in range(3):
            got = vo.pair_distances_simd(rows, q, pr, np.full(40, qi, np.uint32), metric)
            want = np.array([np_simd_distance(q[qi], rows[r], metric) for r in range(40)], np.float32)
            assert np.array_equal(bits(got), bits(want))
        assert np.array_equal(bits(vo.pair_distances_simd(rows, q, pr, None, metric)),
                              bits(vo.pair_distances_simd(rows, q, pr, np.zeros(40, np.uint32), metric)))
    if dim >= 128:  # a different association than the scalar path: some low bits must differ
        scalar = np.array([vo.l2sq(rows[r], q[0]) for r in range(40)], np.float32)
        assert not np.array_equal(bits(scalar), bits(vo.pair_distances_simd(rows, q, pr, None, 0)))
    with pytest.raises(Exception):
        vo.pair_distances_simd(rows, q, np.array([40], np.uint64), None, 1)


def test_golden_simd_vectors_reproduce(vo):
    g = np.load(os.path.join(GOLD, "simd_small.npz"))
    for dim in (300, 67):
        rows = vo.synth(1, 500, dim, kind=1, n_centers=8, center_seed=7, normalize=True)
        q = vo.synth(2, 6, dim, kind=1, n_centers=8, center_seed=7, normalize=True)
        for metric in (0, 1):
            d = vo.pair_distances_simd(rows, q, g[f"pair_row_{dim}"], g[f"pair_query_{dim}"], metric)
            assert np.array_equal(bits(d), g[f"d_{dim}_m{metric}_bits"])


def test_golden_vectors_reproduce(vo):
    g = np.load(os.path.join(GOLD, "c1_small.npz"))
    rows = vo.synth(int(g["seed_data"]), int(g["n"]), int(g["dim"]), kind=1, n_centers=int(g["n_centers"]),
                    center_seed=int(g["seed_centers"]))
    q = vo.synth(int(g["seed_query"]), int(g["nq"]), int(g["dim"]), kind=1, n_centers=int(g["n_centers"]),
                 center_seed=int(g["seed_centers"]))
    assert np.array_equal(bits(rows[:8]), g["rows_head_bits"])
    C = int(g["C"])
    init = vo.init_rows(int(g["seed_init"]), 1, C, int(g["n"]))
    assert np.array_equal(init[0], g["init_rows"])
    cents, assign, cost, _ = vo.ivf_build_index(rows, C, 1, int(g["max_iter"]), init)
    assert np.array_equal(assign, g["assign"]) and np.array_equal(bits(cents), g["cents_bits"])
    assert bits(cost) == g["cost_bits"]
    off, lr = vo.ivf_lists(assign, C)
    for name, nprobe in (("ref", 0), ("np4", 4)):
        ids, d, cnt = vo.ivf_search(rows, cents, off, lr, q, int(g["k"]), nprobe=nprobe)
        assert np.array_equal(ids, g[f"ids_{name}"]) and np.array_equal(bits(d), g[f"d_{name}_bits"])
    ids, d, _ = vo.exhaustive(rows, q, int(g["k"]))
    assert np.array_equal(ids, g["ids_flat"]) and np.array_equal(bits(d), g["d_flat_bits"])
    planes, consts = g["planes_bits"].view(np.float32), g["consts_bits"].view(np.float32)
    assert np.array_equal(np.packbits(vo.lsh_hash(rows, planes, consts)), g["hash_packed"])
    L = vo.LSH(rows, None, int(g["trees"]), int(g["max_size"]), int(g["seed_lsh"]))
    ids, d, _ = L.search(q, int(g["k"]))
    assert np.array_equal(ids, g["ids_lsh"]) and np.array_equal(bits(d), g["d_lsh_bits"])
