"""ctypes binding of the CPU oracle (oracle/vers_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  Nothing under vers_b200/ may import this package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libvers_oracle.so")

OK, ERR_PANIC, ERR_ARG, ERR_NOMEM = 0, -1, -2, -3


class OraclePanic(RuntimeError):
    """The reference would panic on these inputs (index out of bounds / unwrap on None)."""


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "vers_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "vers_synth.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr)
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


_lib = None

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
u64, u32, i32, f32 = C.c_uint64, C.c_uint32, C.c_int, C.c_float


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    sig = {
        "vo_dot": (f32, [_f32p, _f32p, u32]),
        "vo_l2sq": (f32, [_f32p, _f32p, u32]),
        "vo_cosine_distance": (f32, [_f32p, _f32p, u32]),
        "vo_normalize_rows": (None, [_f32p, u64, u32, u32]),
        "vo_synth": (None, [u64, u64, u32, u32, u64, u64, u32, u32, i32, _f32p]),
        "vo_synth_rows": (None, [u64, u64, u32, u32, _u64p, u64, u32, u32, i32, _f32p]),
        "vo_assign": (i32, [_f32p, u64, u32, u32, _f32p, u32, u32, _u64p]),
        "vo_partial_sums": (None, [_f32p, u64, u32, u32, _u64p, u32, _f32p, _u64p]),
        "vo_finalize_centroids": (None, [_f32p, _u64p, u32, u32, _f32p]),
        "vo_update": (i32, [_f32p, u64, u32, u32, _u64p, u32, _f32p, _u64p]),
        "vo_update_sharded": (i32, [_f32p, u64, u32, u32, _u64p, u32, u32, _f32p, _u64p]),
        "vo_kmeans_cost": (f32, [_f32p, u64, u32, u32, _f32p, u32, _u64p]),
        "vo_kmeans_fit": (i32, [_f32p, u64, u32, u32, _u64p, u32, u32, u32, _f32p, _u64p, C.POINTER(u32)]),
        "vo_ivf_build_index": (
            i32,
            [_f32p, u64, u32, u32, u32, u32, u32, _u64p, _f32p, _u64p, C.POINTER(f32), C.POINTER(u32)],
        ),
        "vo_ivf_lists": (None, [_u64p, u64, u32, _u64p, _u64p]),
        "vo_ivf_search_batch": (
            i32,
            [_f32p, u32, u32, _f32p, u32, u32, _u64p, _u64p, _f32p, u32, u32, u32, u32, _u64p, _f32p, _u32p],
        ),
        "vo_nearest_centroid": (i32, [_f32p, u32, u32, u32, _f32p, C.POINTER(u32)]),
        "vo_exhaustive_batch": (i32, [_f32p, u64, u32, u32, _f32p, u32, u32, u32, u32, u64, _u64p, _f32p, _u32p]),
        "vo_simd_distance": (f32, [_f32p, _f32p, u32, u32]),
        "vo_pair_distances_simd": (i32, [_f32p, u64, u32, u32, _f32p, u32, u32, C.c_void_p, _u64p, u64, u32, _f32p]),
        "vo_lsh_hash": (None, [_f32p, u64, u32, u32, _f32p, u32, u32, _f32p, _u8p]),
        "vo_lsh_make_plane": (None, [_f32p, _f32p, u32, _f32p, C.POINTER(f32)]),
        "vo_lsh_build": (C.c_void_p, [_f32p, u64, u32, u32, C.c_void_p, u32, u32, u64]),
        "vo_lsh_error": (i32, [C.c_void_p]),
        "vo_lsh_num_values": (u64, [C.c_void_p]),
        "vo_lsh_tree_nodes": (u32, [C.c_void_p, u32]),
        "vo_lsh_search_batch": (i32, [C.c_void_p, _f32p, u32, u32, u32, _u64p, _f32p, _u32p]),
        "vo_lsh_candidates": (i32, [C.c_void_p, _f32p, u32, _u32p, u32, C.POINTER(u32)]),
        "vo_lsh_add": (i32, [C.c_void_p, _f32p, u64]),
        "vo_lsh_flatten": (
            None,
            [C.c_void_p, u32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
             C.POINTER(u32), C.POINTER(u32), C.POINTER(u64)],
        ),
        "vo_lsh_free": (None, [C.c_void_p]),
        "vo_num_threads": (i32, []),
        "vo_set_threads": (None, [i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _chk(rc: int, what: str):
    if rc == ERR_PANIC:
        raise OraclePanic(f"{what}: the reference would panic on these inputs")
    if rc != OK:
        raise RuntimeError(f"{what}: oracle error {rc}")


def _rows(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2
    return a


# ---------------------------------------------------------------- primitives
def dot(a, b) -> np.float32:
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return np.float32(lib().vo_dot(a, b, a.shape[0]))


def l2sq(a, b) -> np.float32:
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return np.float32(lib().vo_l2sq(a, b, a.shape[0]))


def cosine_distance(a, b) -> np.float32:
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return np.float32(lib().vo_cosine_distance(a, b, a.shape[0]))


def normalize_rows(rows) -> np.ndarray:
    out = _rows(rows).copy()
    lib().vo_normalize_rows(out, out.shape[0], out.shape[1], out.shape[1])
    return out


def synth(seed, n, dim, kind=0, n_centers=1, center_seed=0, row0=0, normalize=True, stride=None) -> np.ndarray:
    stride = stride or dim
    out = np.empty((n, stride), np.float32)
    lib().vo_synth(seed, center_seed, kind, n_centers, row0, n, dim, stride, int(normalize), out)
    return out if stride != dim else out


def synth_rows(seed, row_ids, dim, kind=0, n_centers=1, center_seed=0, normalize=True) -> np.ndarray:
    """rows `row_ids` of the synthetic matrix (same bits as synth(...)[row_ids])"""
    ids = np.ascontiguousarray(row_ids, np.uint64)
    out = np.empty((ids.shape[0], dim), np.float32)
    lib().vo_synth_rows(seed, center_seed, kind, n_centers, ids, ids.shape[0], dim, dim, int(normalize), out)
    return out


def init_rows(seed, attempts, C_, n) -> np.ndarray:
    """k-means initial centroid rows, see include/vers_synth.h (vers_synth_init_row)."""
    M = (1 << 64) - 1

    def sm(x):
        x = (x + 0x9E3779B97F4A7C15) & M
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M
        return x ^ (x >> 31)

    s = sm(seed)
    out = np.empty((attempts, C_), np.uint64)
    for t in range(attempts):
        for j in range(C_):
            out[t, j] = sm((s + (t << 32) + j) & M) % n
    return out


# ---------------------------------------------------------------- k-means / IVFFlat
def assign(rows, cents) -> np.ndarray:
    rows, cents = _rows(rows), _rows(cents)
    out = np.empty(rows.shape[0], np.uint64)
    _chk(lib().vo_assign(rows, rows.shape[0], rows.shape[1], rows.shape[1], cents, cents.shape[0], cents.shape[1], out),
         "assign")
    return out


def update(rows, assignments, C_, shards=0):
    rows = _rows(rows)
    a = np.ascontiguousarray(assignments, np.uint64)
    cents = np.empty((C_, rows.shape[1]), np.float32)
    counts = np.empty(C_, np.uint64)
    if shards and shards > 1:
        _chk(lib().vo_update_sharded(rows, rows.shape[0], rows.shape[1], rows.shape[1], a, C_, shards, cents, counts),
             "update_sharded")
    else:
        _chk(lib().vo_update(rows, rows.shape[0], rows.shape[1], rows.shape[1], a, C_, cents, counts), "update")
    return cents, counts


def partial_sums(rows, assignments, C_, sums, counts):
    """continue the running per-cluster sums/counts IN PLACE over these rows in row order (ivfflat.rs:52-55)"""
    rows = _rows(rows)
    a = np.ascontiguousarray(assignments, np.uint64)
    assert sums.dtype == np.float32 and sums.shape == (C_, rows.shape[1]) and sums.flags.c_contiguous
    assert counts.dtype == np.uint64 and counts.flags.c_contiguous
    lib().vo_partial_sums(rows, rows.shape[0], rows.shape[1], rows.shape[1], a, C_, sums, counts)


def finalize_centroids(sums, counts):
    out = np.empty_like(sums)
    lib().vo_finalize_centroids(sums, counts, sums.shape[0], sums.shape[1], out)
    return out


def kmeans_cost(rows, cents, assignments) -> np.float32:
    rows, cents = _rows(rows), _rows(cents)
    a = np.ascontiguousarray(assignments, np.uint64)
    return np.float32(lib().vo_kmeans_cost(rows, rows.shape[0], rows.shape[1], rows.shape[1], cents, cents.shape[1], a))


def kmeans_fit(rows, init, max_iter, shards=0):
    rows = _rows(rows)
    init = np.ascontiguousarray(init, np.uint64)
    C_ = init.shape[0]
    cents = np.empty((C_, rows.shape[1]), np.float32)
    a = np.empty(rows.shape[0], np.uint64)
    iters = u32(0)
    _chk(lib().vo_kmeans_fit(rows, rows.shape[0], rows.shape[1], rows.shape[1], init, C_, max_iter, shards, cents, a,
                             C.byref(iters)), "kmeans_fit")
    return cents, a, iters.value


def ivf_build_index(rows, C_, num_attempts, max_iter, init):
    rows = _rows(rows)
    init = np.ascontiguousarray(init, np.uint64).reshape(num_attempts, C_)
    cents = np.zeros((C_, rows.shape[1]), np.float32)
    a = np.zeros(rows.shape[0], np.uint64)
    cost, best = f32(0), u32(0)
    _chk(lib().vo_ivf_build_index(rows, rows.shape[0], rows.shape[1], rows.shape[1], C_, num_attempts, max_iter, init,
                                  cents, a, C.byref(cost), C.byref(best)), "ivf_build_index")
    return cents, a, np.float32(cost.value), best.value


def ivf_lists(assignments, C_):
    a = np.ascontiguousarray(assignments, np.uint64)
    off = np.empty(C_ + 1, np.uint64)
    lr = np.empty(a.shape[0], np.uint64)
    lib().vo_ivf_lists(a, a.shape[0], C_, off, lr)
    return off, lr


def ivf_search(rows, cents, offsets, list_rows, queries, k, nprobe=0):
    """nprobe=0: reference semantics (nearest list + spill, concatenated); nprobe>=1: global top-k extension."""
    rows, cents, queries = _rows(rows), _rows(cents), _rows(queries)
    nq = queries.shape[0]
    ids = np.full((nq, max(k, 1)), np.iinfo(np.uint64).max, np.uint64)
    d = np.full((nq, max(k, 1)), np.inf, np.float32)
    cnt = np.zeros(nq, np.uint32)
    rc = lib().vo_ivf_search_batch(rows, rows.shape[1], rows.shape[1], cents, cents.shape[0], cents.shape[1],
                                   np.ascontiguousarray(offsets, np.uint64), np.ascontiguousarray(list_rows, np.uint64),
                                   queries, nq, queries.shape[1], k, nprobe, ids, d, cnt)
    _chk(rc, "ivf_search")
    return ids[:, :k], d[:, :k], cnt


def nearest_centroid(cents, x) -> int:
    cents = _rows(cents)
    x = np.ascontiguousarray(x, np.float32)
    out = u32(0)
    _chk(lib().vo_nearest_centroid(cents, cents.shape[0], cents.shape[1], cents.shape[1], x, C.byref(out)), "nearest")
    return out.value


def exhaustive(rows, queries, k, metric=0, id_base=0):
    rows, queries = _rows(rows), _rows(queries)
    nq = queries.shape[0]
    ids = np.full((nq, max(k, 1)), np.iinfo(np.uint64).max, np.uint64)
    d = np.full((nq, max(k, 1)), np.inf, np.float32)
    cnt = np.zeros(nq, np.uint32)
    _chk(lib().vo_exhaustive_batch(rows, rows.shape[0], rows.shape[1], rows.shape[1], queries, nq, queries.shape[1], k,
                                   metric, id_base, ids, d, cnt), "exhaustive")
    return ids[:, :k], d[:, :k], cnt


def pair_distances_simd(rows, queries, pair_row, pair_query=None, metric=1) -> np.ndarray:
    """HNSW's distances (base.rs:158-294, SIMD association) for a batch of (query, row) pairs; metric 1 = cosine
    distance (what hnsw.rs calls), 0 = squared euclidean"""
    rows, queries = _rows(rows), _rows(queries)
    pair_row = np.ascontiguousarray(pair_row, np.uint64)
    pq = None if pair_query is None else np.ascontiguousarray(pair_query, np.uint32)
    out = np.empty(pair_row.shape[0], np.float32)
    _chk(lib().vo_pair_distances_simd(rows, rows.shape[0], rows.shape[1], rows.shape[1], queries, queries.shape[0],
                                      queries.shape[1], None if pq is None else pq.ctypes.data, pair_row,
                                      pair_row.shape[0], metric, out), "pair_distances_simd")
    return out


# ---------------------------------------------------------------- LSH
def lsh_hash(rows, planes, consts) -> np.ndarray:
    rows, planes = _rows(rows), _rows(planes)
    consts = np.ascontiguousarray(consts, np.float32)
    bits = np.empty((rows.shape[0], planes.shape[0]), np.uint8)
    lib().vo_lsh_hash(rows, rows.shape[0], rows.shape[1], rows.shape[1], planes, planes.shape[0], planes.shape[1],
                      consts, bits)
    return bits


def lsh_make_plane(va, vb):
    va = np.ascontiguousarray(va, np.float32)
    vb = np.ascontiguousarray(vb, np.float32)
    coef = np.empty_like(va)
    c = f32(0)
    lib().vo_lsh_make_plane(va, vb, va.shape[0], coef, C.byref(c))
    return coef, np.float32(c.value)


class LSH:
    """Oracle ANNIndex (lsh.rs:47-283) with the injected sample pairs of include/vers_synth.h."""

    def __init__(self, rows, ids, num_trees, max_size, seed):
        rows = _rows(rows)
        self.dim = rows.shape[1]
        idp = None
        if ids is not None:
            self._ids = np.ascontiguousarray(ids, np.uint64)
            idp = self._ids.ctypes.data_as(C.c_void_p)
        self.h = lib().vo_lsh_build(rows, rows.shape[0], rows.shape[1], rows.shape[1], idp, num_trees, max_size, seed)
        self.num_trees = num_trees
        _chk(lib().vo_lsh_error(self.h), "lsh_build")

    def __del__(self):
        if getattr(self, "h", None):
            lib().vo_lsh_free(self.h)
            self.h = None

    @property
    def num_values(self) -> int:
        return lib().vo_lsh_num_values(self.h)

    def search(self, queries, k):
        queries = _rows(queries)
        nq = queries.shape[0]
        ids = np.full((nq, max(k, 1)), np.iinfo(np.uint64).max, np.uint64)
        d = np.full((nq, max(k, 1)), np.inf, np.float32)
        cnt = np.zeros(nq, np.uint32)
        _chk(lib().vo_lsh_search_batch(self.h, queries, nq, queries.shape[1], k, ids, d, cnt), "lsh_search")
        return ids[:, :k], d[:, :k], cnt

    def candidates(self, q, k) -> np.ndarray:
        q = np.ascontiguousarray(q, np.float32)
        cap = self.num_trees * max(k, 1) * 4 + 64
        out = np.empty(cap, np.uint32)
        n = u32(0)
        lib().vo_lsh_candidates(self.h, q, k, out, cap, C.byref(n))
        assert n.value <= cap
        return out[: n.value].copy()

    def add(self, x, vec_id):
        x = np.ascontiguousarray(x, np.float32)
        _chk(lib().vo_lsh_add(self.h, x, vec_id), "lsh_add")

    def flatten(self, tree):
        nn, ni, nit = u32(0), u32(0), u64(0)
        lib().vo_lsh_flatten(self.h, tree, None, None, None, None, None, C.byref(nn), C.byref(ni), C.byref(nit))
        kind = np.empty(nn.value, np.uint8)
        leaf_len = np.empty(nn.value, np.uint32)
        planes = np.empty((ni.value, self.dim), np.float32)
        consts = np.empty(ni.value, np.float32)
        items = np.empty(nit.value, np.uint32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        lib().vo_lsh_flatten(self.h, tree, p(kind), p(leaf_len), p(planes), p(consts), p(items), C.byref(nn),
                             C.byref(ni), C.byref(nit))
        return dict(kind=kind, leaf_len=leaf_len, planes=planes, consts=consts, items=items)


def num_threads() -> int:
    return lib().vo_num_threads()


def set_threads(t: int):
    lib().vo_set_threads(t)
