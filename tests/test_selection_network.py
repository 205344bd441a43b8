"""CPU model of the warp-wide selection networks used by the tensor-core kernels (no GPU needed).

The CUDA code keeps the 32 smallest (key, position) entries of a stream in a sorted list spread over the 32 lanes of a
warp (csrc/ivf_tc.cuh: sel_flush; csrc/ivf.cu: cand_merge_kernel's fold):
  * a batch of <= 32 queued entries is sorted DESCENDING by a bitonic sorting network (partner = lane ^ j),
  * the lane-wise minimum of the ascending list and the descending batch is the smaller half of their union, as a
    bitonic sequence, which a 5-step bitonic merge sorts ascending,
  * cand_merge folds sorted runs into R registers per lane by cascading (min, max) merge-splits.
This file restates those networks lane by lane in numpy (same partner/keep-min rules as the device code) and checks
them against a plain sort on random inputs with ties, infinities and short batches — the arithmetic-free part of the
kernels' correctness argument."""
import numpy as np
import pytest

INF = np.float32(np.inf)
EMPTY = 0xFFFF


def less(d0, r0, d1, r1):
    return (d0 < d1) | ((d0 == d1) & (r0 < r1))


def cmpx(d, r, j, keep_min):
    """one compare-exchange step across the 32 lanes: partner = lane ^ j (sel_cmpx / cm_merge_asc)"""
    lane = np.arange(32)
    od, orr = d[lane ^ j], r[lane ^ j]
    other_less, self_less = less(od, orr, d, r), less(d, r, od, orr)
    take = np.where(keep_min, other_less, self_less)
    return np.where(take, od, d), np.where(take, orr, r)


def sort_desc(d, r):
    lane = np.arange(32)
    k = 2
    while k <= 32:
        j = k >> 1
        while j > 0:
            desc_block = ((lane & k) == 0) | (k == 32)
            lower = (lane & j) == 0
            d, r = cmpx(d, r, j, np.where(desc_block, ~lower, lower))
            j >>= 1
        k <<= 1
    return d, r


def merge_asc(d, r):
    lane = np.arange(32)
    j = 16
    while j > 0:
        d, r = cmpx(d, r, j, (lane & j) == 0)
        j >>= 1
    return d, r


def sel_flush(list_d, list_r, queue_d, queue_r):
    """csrc/ivf_tc.cuh sel_flush: fold c <= 32 queued entries into the sorted 32-entry list"""
    c = len(queue_d)
    d = np.full(32, INF, np.float32)
    r = np.full(32, EMPTY, np.uint32)
    d[:c], r[:c] = queue_d, queue_r
    d, r = sort_desc(d, r)
    take = less(d, r, list_d, list_r)
    ld, lr = np.where(take, d, list_d), np.where(take, r, list_r)
    return merge_asc(ld, lr)


def as_sorted_pairs(d, r):
    return sorted(zip(d.tolist(), r.tolist()))


@pytest.mark.parametrize("seed", range(20))
def test_bitonic_sort_descending(seed):
    rng = np.random.default_rng(seed)
    d = rng.integers(0, 6, 32).astype(np.float32)  # many ties
    d[rng.integers(0, 32, 4)] = INF
    r = rng.permutation(32).astype(np.uint32)
    sd, sr = sort_desc(d.copy(), r.copy())
    assert list(zip(sd.tolist(), sr.tolist())) == sorted(zip(d.tolist(), r.tolist()), reverse=True)


@pytest.mark.parametrize("seed", range(30))
def test_sel_flush_keeps_the_32_smallest(seed):
    rng = np.random.default_rng(100 + seed)
    pool_d = rng.integers(0, 40, 200).astype(np.float32) / np.float32(8)
    pool_r = rng.permutation(4096)[:200].astype(np.uint32)
    list_d = np.full(32, INF, np.float32)
    list_r = np.full(32, EMPTY, np.uint32)
    seen = []
    at = 0
    while at < 200:
        c = int(rng.integers(1, 33))  # batches of 1..32 entries, like a queue flushed early or at the end of an item
        qd, qr = pool_d[at:at + c], pool_r[at:at + c]
        seen += list(zip(qd.tolist(), qr.tolist()))
        list_d, list_r = sel_flush(list_d, list_r, qd, qr)
        want = sorted(seen)[:32]
        got = [(a, b) for a, b in zip(list_d.tolist(), list_r.tolist()) if b != EMPTY]
        assert got == want
        at += c


@pytest.mark.parametrize("R", [1, 2, 4])
@pytest.mark.parametrize("seed", range(6))
def test_cand_merge_fold_keeps_the_top_32R(R, seed):
    """csrc/ivf.cu cand_merge_kernel: sorted runs of 32 folded into R sorted registers per lane by merge-splits"""
    rng = np.random.default_rng(1000 + 10 * R + seed)
    ad = [np.full(32, INF, np.float32) for _ in range(R)]
    ap = [np.full(32, 0xFFFFFFFF, np.uint32) for _ in range(R)]
    everything = []
    lane = np.arange(32)
    for _ in range(12):
        n = int(rng.integers(0, 33))
        d = np.full(32, INF, np.float32)
        p = np.full(32, 0xFFFFFFFF, np.uint32)
        vals = sorted(zip((rng.integers(0, 50, n) / 4).astype(np.float32).tolist(),
                          rng.choice(1 << 20, n, replace=False).astype(np.uint32).tolist()))
        for i, (a, b) in enumerate(vals):
            d[i], p[i] = a, b
        everything += vals
        if n == 0 or not less(d[0], p[0], ad[R - 1][31], ap[R - 1][31]):
            continue  # the run's head cannot enter: skipped like on the device
        for rr in range(R):
            rd, rp = d[31 - lane], p[31 - lane]
            take = less(rd, rp, ad[rr], ap[rr])
            lo_d, lo_p = np.where(take, rd, ad[rr]), np.where(take, rp, ap[rr])
            hi_d, hi_p = np.where(take, ad[rr], rd), np.where(take, ap[rr], rp)
            ad[rr], ap[rr] = merge_asc(lo_d, lo_p)
            if rr + 1 < R:
                d, p = merge_asc(hi_d, hi_p)
    got = [(a, b) for reg_d, reg_p in zip(ad, ap) for a, b in zip(reg_d.tolist(), reg_p.tolist()) if b != 0xFFFFFFFF]
    assert got == sorted(everything)[:32 * R]


def test_threshold_encoding_is_order_preserving():
    """tau_encode / tau_decode (csrc/ivf_tc.cuh): float -> uint32 that sorts like the float, +inf = 0xff800000"""
    def enc(f):
        b = np.float32(f).view(np.uint32)
        return np.uint32(~b) if b & np.uint32(0x80000000) else np.uint32(b | np.uint32(0x80000000))

    def dec(u):
        u = np.uint32(u)
        return (np.uint32(u & np.uint32(0x7FFFFFFF)) if u & np.uint32(0x80000000) else np.uint32(~u)).view(np.float32)

    vals = np.array([-np.inf, -3.5, -1e-30, -0.0, 0.0, 1e-38, 0.25, 1.0, 7e30, np.inf], np.float32)
    codes = [int(enc(v)) for v in vals]
    assert codes == sorted(codes)
    assert int(enc(np.float32(np.inf))) == 0xFF800000
    for v in vals:
        assert dec(enc(v)).view(np.uint32) == np.float32(v).view(np.uint32)


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("ties", [False, True])
def test_shared_bound_of_four_slots_never_drops_a_top32_row(seed, ties):
    """The per-query bound shared by the work items of a candidate scan (csrc/ivf_tc.cuh TcScanParams::qtau,
    csrc/flat_tc.cuh): 4 slots, a list publishes its 8th key (atomicMin) into the slot of its class, rows are pruned
    against the LARGEST slot and against the own list's 32nd key.  Lists of different classes cover disjoint rows, so
    32 distinct rows lie at or below the bound.  Model: items run in a random interleaving with stale reads of the
    slots; whatever the timing, the union of the partial lists must contain the true 32 smallest (key, position)
    entries, and every dropped row must have key >= the 32nd merged key (the bound cand_merge hands the certificate)."""
    rng = np.random.default_rng(100 + seed)
    n_items, rows_per_item, tile = 12, 512, 128
    n = n_items * rows_per_item
    keys = rng.standard_normal(n).astype(np.float32)
    if ties:
        keys = np.round(keys * 4) / np.float32(4)  # many equal keys, also at the bound
    slots = np.full(4, INF, np.float32)
    # list state per (item, lane group): sorted (key, pos) entries, at most 32
    lists = {(it, g): [] for it in range(n_items) for g in range(4)}
    published = {(it, g): INF for it in range(n_items) for g in range(4)}
    dropped = []
    # every item is a sequence of tiles; interleave the tiles of all items at random (CTAs progress independently)
    cursor = [0] * n_items
    stale = {it: INF for it in range(n_items)}  # the bound an item last read (refreshed once per tile, may be old)
    pending = [it for it in range(n_items) for _ in range(rows_per_item // tile)]
    rng.shuffle(pending)
    for it in pending:
        t0 = cursor[it]
        cursor[it] += tile
        if rng.random() < 0.7:  # a refresh may be skipped: stale values are larger, hence safe
            stale[it] = min(stale[it], float(slots.max()))
        for g in range(4):  # lane group g owns rows g*32 .. g*32+31 of the tile
            L = lists[(it, g)]
            for r in range(t0 + g * 32, t0 + g * 32 + 32):
                pos = it * rows_per_item + r
                own32 = L[31][0] if len(L) == 32 else INF
                tau = min(stale[it], own32)
                if keys[pos] <= tau:
                    L.append((float(keys[pos]), pos))
                    L.sort()
                    if len(L) > 32:
                        dropped.append(L.pop())  # evicted by 32 smaller entries of its own list
                else:
                    dropped.append((float(keys[pos]), pos))
            if len(L) >= 8 and L[7][0] < published[(it, g)]:  # flush: publish the 8th key into slot g
                published[(it, g)] = L[7][0]
                slots[g] = min(slots[g], np.float32(L[7][0]))
    merged = sorted(e for L in lists.values() for e in L)
    truth = sorted((float(k), i) for i, k in enumerate(keys))
    assert merged[:32] == truth[:32]
    bound = merged[31][0]
    assert all(k >= bound for k, _ in dropped)
    assert float(slots.max()) >= truth[31][0]  # the bound itself never undercuts the true 32nd key
