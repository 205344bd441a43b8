"""Host-side mirror of the reference's index interface over the C ABI (include/vers_device.h).

Mirrors, with the same names and argument meaning:
  - ``trait Index<N>``            indexes/base.rs:27-59   add / search_approximate / save_index / load_index
  - ``IVFFlatIndex::build_index`` indexes/ivfflat.rs:102-136
  - ``ANNIndex::build_index``     indexes/lsh.rs:132-161
  - ``utils::search_exhaustive``  utils.rs:68-82 (FlatIndex / search_exhaustive)
Everything numeric happens on the GPU behind the C ABI; this file only moves buffers and keeps the host copy of
the index fields that the reference serialises (values / centroids / assignments / ids).

The reference draws its random choices from ``rand::thread_rng()``; here they are injected (``init_rows`` /
``seed``) so results are reproducible and comparable with the oracle.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _abi
from ._abi import VersError, VersPanic, check, lib, ptr

U64_MAX = np.iinfo(np.uint64).max


def _rows(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim == 1:
        a = a[None, :]
    if a.ndim != 2:
        raise ValueError("expected a [n, dim] float32 array")
    return a


class Context:
    """One per GPU (vers_ctx)."""

    def __init__(self, device: int = 0):
        self.h = C.c_void_p()
        check(lib().vers_ctx_create(device, C.byref(self.h)))
        self.device = device

    def close(self):
        if getattr(self, "h", None) and self.h:
            lib().vers_ctx_destroy(self.h)
            self.h = None

    __del__ = close

    def set_stream(self, cuda_stream: int):
        check(lib().vers_ctx_set_stream(self.h, C.c_void_p(cuda_stream)))

    def sync(self):
        check(lib().vers_ctx_sync(self.h))

    @property
    def launch_count(self) -> int:
        out = C.c_uint64(0)
        check(lib().vers_ctx_launch_count(self.h, C.byref(out)))
        return out.value

    def enable_timing(self, on: bool = True):
        check(lib().vers_ctx_enable_timing(self.h, int(on)))

    def kernel_ms(self, family: int) -> Tuple[float, int]:
        """(total device ms, launches) of one kernel family since enable_timing(True)"""
        ms, n = C.c_float(0), C.c_uint64(0)
        check(lib().vers_ctx_kernel_ms(self.h, family, C.byref(ms), C.byref(n)))
        return ms.value, n.value


_default_ctx: dict = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class Dataset:
    """``Vec<Vector<N>>`` resident in HBM (vers_dataset)."""

    def __init__(self, ctx: Context, handle: C.c_void_p):
        self.ctx = ctx
        self.h = handle
        n, dim, ld, base = C.c_uint64(), C.c_uint32(), C.c_uint32(), C.c_uint64()
        check(lib().vers_dataset_info(self.h, C.byref(n), C.byref(dim), C.byref(ld), C.byref(base)))
        self.n, self.dim, self.ld, self.id_base = n.value, dim.value, ld.value, base.value

    @classmethod
    def upload(cls, ctx: Context, rows, id_base: int = 0) -> "Dataset":
        rows = _rows(rows)
        h = C.c_void_p()
        check(lib().vers_dataset_upload(ctx.h, ptr(rows), rows.shape[0], rows.shape[1], rows.shape[1], id_base,
                                        C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def upload_strided(cls, ctx: Context, buf: np.ndarray, n: int, dim: int, stride: int, id_base: int = 0) -> "Dataset":
        """rows laid out like the reference's Vec<Vector<N>>: stride = size_of::<Vector<N>>() / 4 floats"""
        buf = np.ascontiguousarray(buf, np.float32)
        h = C.c_void_p()
        check(lib().vers_dataset_upload(ctx.h, ptr(buf), n, dim, stride, id_base, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def synth(cls, ctx: Context, seed: int, n: int, dim: int, kind: int = 0, n_centers: int = 1, center_seed: int = 0,
              row0: int = 0, normalize: bool = True) -> "Dataset":
        h = C.c_void_p()
        check(lib().vers_dataset_synth(ctx.h, seed, center_seed, kind, n_centers, row0, n, dim, int(normalize),
                                       C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def wrap_device(cls, ctx: Context, d_rows_ptr: int, n: int, dim: int, id_base: int = 0, keepalive=None) -> "Dataset":
        """non-owning handle over rows already in device memory ([n][round_up(dim, 4)] fp32); ``keepalive`` (e.g. the
        torch tensor that owns the memory) is held for the life of the handle"""
        h = C.c_void_p()
        check(lib().vers_dataset_wrap_device(ctx.h, C.c_void_p(d_rows_ptr), n, dim, id_base, C.byref(h)))
        ds = cls(ctx, h)
        ds._keepalive = keepalive
        return ds

    def normalize(self):
        check(lib().vers_dataset_normalize(self.h))

    def download(self, row0: int = 0, n: Optional[int] = None) -> np.ndarray:
        n = self.n - row0 if n is None else n
        out = np.empty((n, self.dim), np.float32)
        check(lib().vers_dataset_download(self.h, row0, n, ptr(out), self.dim))
        return out

    @property
    def device_ptr(self) -> int:
        p = C.c_void_p()
        check(lib().vers_dataset_device_ptr(self.h, C.byref(p)))
        return p.value

    def set_flat_mode(self, mode: int):
        """exhaustive search: 0 (default) tensor-core candidate path for batches >= 32 queries, 1 exact-order only"""
        check(lib().vers_flat_set_mode(self.h, int(mode)))

    def last_flat_search_stats(self) -> dict:
        out = np.zeros(8, np.uint64)
        check(lib().vers_flat_last_search_stats(self.h, ptr(out)))
        return dict(uncertified_queries=int(out[4]), reranked=int(out[5]),
                    max_candidate_error=float(np.array([out[6]], np.uint64).astype(np.uint32).view(np.float32)[0]))

    def close(self):
        if getattr(self, "h", None) and self.h:
            lib().vers_dataset_free(self.h)
            self.h = None

    __del__ = close


def _result_buffers(nq: int, k: int):
    ids = np.full((nq, max(k, 1)), U64_MAX, np.uint64)
    d = np.full((nq, max(k, 1)), np.inf, np.float32)
    cnt = np.zeros(nq, np.uint32)
    return ids, d, cnt


def _to_pairs(ids: np.ndarray, d: np.ndarray, cnt: int) -> List[Tuple[int, float]]:
    return [(int(ids[i]), float(d[i])) for i in range(cnt)]


# --------------------------------------------------------------------------------------------- exhaustive
def search_exhaustive_batch(ds: Dataset, queries, top_k: int, metric: int = _abi.METRIC_L2SQ):
    q = _rows(queries)
    ids, d, cnt = _result_buffers(q.shape[0], top_k)
    check(lib().vers_flat_search(ds.h, ptr(q), q.shape[0], q.shape[1], top_k, metric, ptr(ids), ptr(d), ptr(cnt)))
    return ids[:, :top_k], d[:, :top_k], cnt


def search_exhaustive(ds: Dataset, query, top_k: int) -> List[Tuple[int, float]]:
    """utils::search_exhaustive(&data, &query, k) (utils.rs:68-82)"""
    ids, d, cnt = search_exhaustive_batch(ds, query, top_k)
    return _to_pairs(ids[0], d[0], int(cnt[0]))


def pair_distances_simd(ds: Dataset, queries, pair_row, pair_query=None, metric: int = _abi.METRIC_COSINE) -> np.ndarray:
    """HNSW distance offload: ``Vector::cosine_similarity_simd`` (base.rs:158-223; metric 1, what hnsw.rs:146/258/273
    call) or ``squared_euclidean_simd`` (base.rs:225-294; metric 0) for a batch of (query, row id) pairs, in the
    reference's SIMD summation order.  ``pair_query`` None: every pair uses query 0 (one node against its neighbours)."""
    q = _rows(queries)
    rows = np.ascontiguousarray(pair_row, np.uint64)
    pq = None if pair_query is None else np.ascontiguousarray(pair_query, np.uint32)
    if pq is not None and pq.shape[0] != rows.shape[0]:
        raise ValueError("pair_query and pair_row differ in length")
    out = np.empty(rows.shape[0], np.float32)
    check(lib().vers_pair_distances_simd(ds.h, ptr(q), q.shape[0], q.shape[1], None if pq is None else ptr(pq), ptr(rows),
                                         rows.shape[0], metric, ptr(out)))
    return out


# --------------------------------------------------------------------------------------------- k-means steps
class KMeans:
    """Device k-means state on one GPU's row shard (vers_kmeans); the single-GPU driver is ``fit``; the
    multi-GPU driver in vers_b200.sharded chains ``sums_step`` across ranks."""

    def __init__(self, ds: Dataset, num_clusters: int):
        self.ds = ds
        self.C = num_clusters
        self.h = C.c_void_p()
        check(lib().vers_kmeans_create(ds.h, num_clusters, C.byref(self.h)))

    def close(self):
        if getattr(self, "h", None) and self.h:
            lib().vers_kmeans_free(self.h)
            self.h = None

    __del__ = close

    def init_from_rows(self, init_rows):
        r = np.ascontiguousarray(init_rows, np.uint64)
        assert r.shape == (self.C,)
        check(lib().vers_kmeans_init_from_rows(self.h, ptr(r)))

    def set_centroids(self, cents):
        c = _rows(cents)
        assert c.shape[0] == self.C
        check(lib().vers_kmeans_set_centroids(self.h, ptr(c), c.shape[1]))

    def centroids(self) -> np.ndarray:
        out = np.empty((self.C, self.ds.dim), np.float32)
        check(lib().vers_kmeans_get_centroids(self.h, ptr(out), self.ds.dim))
        return out

    def assignments(self) -> np.ndarray:
        out = np.empty(self.ds.n, np.uint64)
        check(lib().vers_kmeans_get_assignments(self.h, ptr(out)))
        return out

    def assign_step(self):
        check(lib().vers_kmeans_assign_step(self.h))

    def set_mode(self, mode: int):
        """0 (default): tensor-core candidate argmin + certificate + exact redo; 1: exact order only"""
        check(lib().vers_kmeans_set_mode(self.h, int(mode)))

    @property
    def last_uncertified_rows(self) -> int:
        out = C.c_uint64(0)
        check(lib().vers_kmeans_last_assign_stats(self.h, C.byref(out)))
        return out.value

    def sums_step_dev(self, d_sums_ptr: int, d_counts_ptr: int):
        check(lib().vers_kmeans_sums_step_dev(self.h, C.c_void_p(d_sums_ptr), C.c_void_p(d_counts_ptr)))

    def finalize_step_dev(self, d_sums_ptr: int, d_counts_ptr: int) -> bool:
        ch = C.c_uint32(0)
        check(lib().vers_kmeans_finalize_step_dev(self.h, C.c_void_p(d_sums_ptr), C.c_void_p(d_counts_ptr),
                                                  C.byref(ch)))
        return bool(ch.value)

    def cost_step(self, cost_in: float = 0.0) -> np.float32:
        c = C.c_float(cost_in)
        check(lib().vers_kmeans_cost_step(self.h, C.byref(c)))
        return np.float32(c.value)

    def fit(self, max_iterations: int) -> int:
        it = C.c_uint32(0)
        check(lib().vers_kmeans_fit(self.h, max_iterations, C.byref(it)))
        return it.value


def assign_to_clusters(ds: Dataset, centroids) -> np.ndarray:
    """IVFFlatIndex::assign_to_clusters (ivfflat.rs:29-46)"""
    c = _rows(centroids)
    out = np.empty(ds.n, np.uint64)
    check(lib().vers_kmeans_assign(ds.h, ptr(c), c.shape[0], c.shape[1], ptr(out)))
    return out


def update_centroids(ds: Dataset, assignments, k: int):
    """IVFFlatIndex::update_centroids (ivfflat.rs:47-71) -> (centroids [k, dim], counts [k])"""
    a = np.ascontiguousarray(assignments, np.uint64)
    cents = np.empty((k, ds.dim), np.float32)
    counts = np.empty(k, np.uint64)
    check(lib().vers_kmeans_update(ds.h, ptr(a), k, ptr(cents), ptr(counts)))
    return cents, counts


# --------------------------------------------------------------------------------------------- IVFFlat
class IVFFlatIndex:
    """IVFFlatIndex<N> (indexes/ivfflat.rs:8-15): {num_centroids, values, centroids, assignments, ids}."""

    def __init__(self, ctx: Context, handle: C.c_void_p, values: Optional[np.ndarray], ds: Optional[Dataset] = None):
        self.ctx = ctx
        self.h = handle
        self._ds = ds  # keeps the row-major device copy alive only as long as the caller wants it
        n, dim, nc, cost, att = C.c_uint64(), C.c_uint32(), C.c_uint32(), C.c_float(), C.c_uint32()
        check(lib().vers_ivf_info(self.h, C.byref(n), C.byref(dim), C.byref(nc), C.byref(cost), C.byref(att)))
        self.dim, self.num_centroids = dim.value, nc.value
        self.best_cost, self.best_attempt = np.float32(cost.value), att.value
        self.values = values  # host copy (Vec<Vector<N>>) kept for save_index, may be None for synthetic shards

    # -- construction
    @classmethod
    def build_index(cls, num_clusters: int, num_attempts: int, max_iterations: int, vectors, *,
                    init_rows: Optional[np.ndarray] = None, seed: int = 3, ctx: Optional[Context] = None,
                    dataset: Optional[Dataset] = None) -> "IVFFlatIndex":
        """IVFFlatIndex::build_index(num_clusters, num_attempts, max_iterations, &vectors) (ivfflat.rs:102-136).
        ``vectors``: [n, dim] array (or None with ``dataset=`` for rows already on the device).
        ``init_rows``: [num_attempts, num_clusters] row numbers standing in for the thread_rng draws."""
        ctx = ctx or (dataset.ctx if dataset is not None else default_context())
        values = None
        if dataset is None:
            values = _rows(vectors).copy()
            dataset = Dataset.upload(ctx, values)
        if init_rows is None:
            init_rows = synth_init_rows(seed, num_attempts, num_clusters, dataset.n)
        init_rows = np.ascontiguousarray(init_rows, np.uint64).reshape(num_attempts, num_clusters)
        h = C.c_void_p()
        check(lib().vers_ivf_build_index(dataset.h, num_clusters, num_attempts, max_iterations, ptr(init_rows),
                                         C.byref(h)))
        return cls(ctx, h, values, dataset)

    @classmethod
    def from_parts(cls, vectors, centroids, assignments=None, *, ctx: Optional[Context] = None,
                   dataset: Optional[Dataset] = None) -> "IVFFlatIndex":
        ctx = ctx or (dataset.ctx if dataset is not None else default_context())
        values = None
        if dataset is None:
            values = _rows(vectors).copy()
            dataset = Dataset.upload(ctx, values)
        c = _rows(centroids)
        a = None if assignments is None else np.ascontiguousarray(assignments, np.uint64)
        h = C.c_void_p()
        check(lib().vers_ivf_from_parts(dataset.h, ptr(c), c.shape[0], c.shape[1], None if a is None else ptr(a),
                                        C.byref(h)))
        return cls(ctx, h, values, dataset)

    @classmethod
    def from_parts_dev(cls, dataset: Dataset, d_centroids_ptr: int, num_clusters: int, d_assign_ptr: int,
                       d_row_ids_ptr: Optional[int] = None) -> "IVFFlatIndex":
        """the index struct from parts resident on the device: centroids [C][ld] fp32, uint32 cluster per row of
        ``dataset`` and (optionally) the uint64 GLOBAL id of every row — how a list-sharded multi-GPU build hands each
        GPU the rows of the lists it owns"""
        h = C.c_void_p()
        check(lib().vers_ivf_from_parts_dev(dataset.h, C.c_void_p(d_centroids_ptr), num_clusters,
                                            C.c_void_p(d_assign_ptr),
                                            C.c_void_p(d_row_ids_ptr) if d_row_ids_ptr else None, C.byref(h)))
        return cls(dataset.ctx, h, None, dataset)

    @classmethod
    def from_kmeans(cls, km: KMeans, values: Optional[np.ndarray] = None) -> "IVFFlatIndex":
        h = C.c_void_p()
        check(lib().vers_ivf_from_kmeans(km.h, C.byref(h)))
        return cls(km.ds.ctx, h, values, km.ds)

    def close(self):
        if getattr(self, "h", None) and self.h:
            lib().vers_ivf_free(self.h)
            self.h = None

    __del__ = close

    # -- fields
    def __len__(self) -> int:
        n = C.c_uint64()
        check(lib().vers_ivf_info(self.h, C.byref(n), None, None, None, None))
        return n.value

    @property
    def centroids(self) -> np.ndarray:
        out = np.empty((self.num_centroids, self.dim), np.float32)
        check(lib().vers_ivf_get_centroids(self.h, ptr(out), self.dim))
        return out

    @property
    def assignments(self) -> np.ndarray:
        out = np.empty(len(self), np.uint64)
        check(lib().vers_ivf_get_assignments(self.h, ptr(out)))
        return out

    @property
    def list_sizes(self) -> np.ndarray:
        out = np.empty(self.num_centroids, np.uint64)
        check(lib().vers_ivf_get_list_sizes(self.h, ptr(out)))
        return out

    def get_list(self, c: int, with_rows: bool = False):
        """ids[c] (and optionally the rows of list c, in list order) straight from the device layout"""
        n = int(self.list_sizes[c])
        ids = np.empty(n, np.uint64)
        rows = np.empty((n, self.dim), np.float32) if with_rows else None
        check(lib().vers_ivf_get_list(self.h, c, ptr(ids), None if rows is None else ptr(rows), self.dim))
        return (ids, rows) if with_rows else ids

    def last_search_stats(self) -> dict:
        out = np.zeros(8, np.uint64)
        check(lib().vers_ivf_last_search_stats(self.h, ptr(out)))
        return dict(distinct_list_rows=int(out[0]), pair_rows=int(out[1]), work_items=int(out[2]),
                    lists_touched=int(out[3]), uncertified_queries=int(out[4]), reranked=int(out[5]),
                    max_candidate_error=float(np.array([out[6]], np.uint64).astype(np.uint32).view(np.float32)[0]),
                    uncertified_probe_queries=int(out[7]) & 0xffffffff, probe_reranked=int(out[7]) >> 32)

    def set_mode(self, mode):
        """0 / False (default): tensor-core candidate pass (TMA + tcgen05 TF32, split precision hi/lo) + exact-order
        rerank + certificate; 1 / True: exact order everywhere; 2: fp32 FMA candidate pass (SIMT) + rerank +
        certificate; 3: tensor-core candidate pass with plain (unsplit) TF32.
        Every mode returns the reference's ids and distance bits."""
        check(lib().vers_ivf_set_mode(self.h, int(mode)))

    @property
    def ids(self) -> List[np.ndarray]:
        """ids: Vec<Vec<usize>> (ivfflat.rs:123-127) rebuilt from the assignments"""
        a = self.assignments
        order = np.argsort(a, kind="stable")
        bounds = np.searchsorted(a[order], np.arange(self.num_centroids + 1))
        return [order[bounds[c]:bounds[c + 1]].astype(np.uint64) for c in range(self.num_centroids)]

    # -- trait Index<N>
    def add(self, embedding, vec_id: int = 0) -> Tuple[int, int]:
        """Index::add (ivfflat.rs:200-213). NOTE: like the reference, ``vec_id`` is ignored and the stored id is
        ``assignments.len()``.  Returns (stored id, cluster)."""
        e = np.ascontiguousarray(embedding, np.float32).reshape(-1)
        assert e.shape[0] == self.dim
        aid, cl = C.c_uint64(), C.c_uint32()
        check(lib().vers_ivf_add(self.h, ptr(e), vec_id, C.byref(aid), C.byref(cl)))
        if self.values is not None:
            self.values = np.vstack([self.values, e[None, :]])
        return aid.value, cl.value

    def add_batch(self, embeddings) -> Tuple[np.ndarray, np.ndarray]:
        """Index::add (ivfflat.rs:200-213) for a batch, in order: exactly len(embeddings) sequential adds (the centroids
        do not move on add).  Returns (stored ids, clusters)."""
        e = _rows(embeddings)
        ids = np.empty(e.shape[0], np.uint64)
        cl = np.empty(e.shape[0], np.uint32)
        check(lib().vers_ivf_add_batch(self.h, ptr(e), e.shape[0], e.shape[1], ptr(ids), ptr(cl)))
        if self.values is not None:
            self.values = np.vstack([self.values, e])
        return ids, cl

    def search_approximate(self, query, top_k: int) -> List[Tuple[int, float]]:
        """Index::search_approximate(query, top_k) -> Vec<(usize, f32)> (ivfflat.rs:153-198), squared L2."""
        ids, d, cnt = self.search_batch(query, top_k, nprobe=0)
        return _to_pairs(ids[0], d[0], int(cnt[0]))

    def search_batch(self, queries, top_k: int, nprobe: int = 0):
        """Batch extension.  nprobe=0: reference semantics per query; nprobe>=1: global top-k over nprobe lists."""
        q = _rows(queries)
        ids, d, cnt = _result_buffers(q.shape[0], top_k)
        check(lib().vers_ivf_search(self.h, ptr(q), q.shape[0], q.shape[1], top_k, nprobe, ptr(ids), ptr(d), ptr(cnt)))
        return ids[:, :top_k], d[:, :top_k], cnt

    def search_batch_dev(self, d_queries_ptr: int, nq: int, top_k: int, nprobe: int, d_ids_ptr: int, d_d_ptr: int,
                         d_cnt_ptr: int):
        check(lib().vers_ivf_search_dev(self.h, C.c_void_p(d_queries_ptr), nq, top_k, nprobe, C.c_void_p(d_ids_ptr),
                                        C.c_void_p(d_d_ptr), C.c_void_p(d_cnt_ptr)))

    def save_index(self, file_path: str):
        """Index::save_index (base.rs:31-43): bincode 1.3 layout of the reference struct."""
        from .bincode import write_ivfflat

        if self.values is None:
            raise VersError(_abi.ERR_ARG, "save_index needs the host copy of the vectors")
        write_ivfflat(file_path, self.num_centroids, self.values, self.centroids, self.assignments)

    @classmethod
    def load_index(cls, file_path: str, ctx: Optional[Context] = None) -> "IVFFlatIndex":
        """Index::load_index (base.rs:45-58) then rebuild the device mirror."""
        from .bincode import read_ivfflat

        num_centroids, values, centroids, assignments, _ids = read_ivfflat(file_path)
        return cls.from_parts(values, centroids, assignments, ctx=ctx)


def synth_init_rows(seed: int, attempts: int, num_clusters: int, n: int) -> np.ndarray:
    """vers_synth_init_row (include/vers_synth.h): the injected stand-in for thread_rng().gen_range(0..n)"""
    M = (1 << 64) - 1

    def sm(x):
        x = (x + 0x9E3779B97F4A7C15) & M
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M
        return x ^ (x >> 31)

    s = sm(seed)
    out = np.empty((attempts, num_clusters), np.uint64)
    for t in range(attempts):
        for j in range(num_clusters):
            out[t, j] = sm((s + (t << 32) + j) & M) % n
    return out


# --------------------------------------------------------------------------------------------- LSH
def lsh_hash(ds: Dataset, planes, consts) -> np.ndarray:
    """Hyperplane::point_is_above (lsh.rs:27-29) for every row x plane -> uint8 [n, P]"""
    p = _rows(planes)
    c = np.ascontiguousarray(consts, np.float32)
    assert c.shape == (p.shape[0],)
    bits = np.empty((ds.n, p.shape[0]), np.uint8)
    check(lib().vers_lsh_hash(ds.h, ptr(p), p.shape[0], p.shape[1], ptr(c), ptr(bits)))
    return bits


class ANNIndex:
    """ANNIndex<N> (indexes/lsh.rs:47-55): {max_node_size, trees, values, ids}"""

    def __init__(self, ctx: Context, handle: C.c_void_p, dim: int):
        self.ctx = ctx
        self.h = handle
        self.dim = dim

    @classmethod
    def build_index(cls, num_trees: int, max_size: int, vectors, vector_ids: Optional[Sequence[int]] = None, *,
                    seed: int = 4, ctx: Optional[Context] = None) -> "ANNIndex":
        """ANNIndex::build_index(num_trees, max_size, &vectors, &vector_ids) (lsh.rs:132-161)"""
        ctx = ctx or default_context()
        v = _rows(vectors)
        ids = None if vector_ids is None else np.ascontiguousarray(vector_ids, np.uint64)
        h = C.c_void_p()
        check(lib().vers_lsh_build_index(ctx.h, ptr(v), v.shape[0], v.shape[1], v.shape[1],
                                         None if ids is None else ptr(ids), num_trees, max_size, seed, C.byref(h)))
        self = cls(ctx, h, v.shape[1])
        self.max_node_size = max_size
        return self

    def close(self):
        if getattr(self, "h", None) and self.h:
            lib().vers_lsh_free(self.h)
            self.h = None

    __del__ = close

    def info(self):
        nv, nt, nn = C.c_uint64(), C.c_uint32(), C.c_uint64()
        check(lib().vers_lsh_info(self.h, C.byref(nv), C.byref(nt), C.byref(nn)))
        return dict(num_values=nv.value, num_trees=nt.value, num_nodes=nn.value)

    def flatten(self, tree: int):
        nn, ni, nit = C.c_uint32(), C.c_uint32(), C.c_uint64()
        check(lib().vers_lsh_flatten(self.h, tree, None, None, None, None, None, C.byref(nn), C.byref(ni),
                                     C.byref(nit)))
        kind = np.empty(nn.value, np.uint8)
        leaf_len = np.empty(nn.value, np.uint32)
        planes = np.empty((ni.value, self.dim), np.float32)
        consts = np.empty(ni.value, np.float32)
        items = np.empty(nit.value, np.uint32)
        check(lib().vers_lsh_flatten(self.h, tree, ptr(kind), ptr(leaf_len), ptr(planes), ptr(consts), ptr(items),
                                     C.byref(nn), C.byref(ni), C.byref(nit)))
        return dict(kind=kind, leaf_len=leaf_len, planes=planes, consts=consts, items=items)

    def values_and_ids(self):
        """`values` / `ids` of the struct (lsh.rs:47-55): the deduplicated rows in stored order (lsh.rs:113-130: first
        occurrence of every bit pattern), rows added later included"""
        n = self.info()["num_values"]
        values = np.empty((n, self.dim), np.float32)
        ids = np.empty(n, np.uint64)
        check(lib().vers_lsh_get_values(self.h, ptr(values), self.dim, ptr(ids)))
        return values, ids

    def save_index(self, file_path: str):
        """Index::save_index (base.rs:31-43): the reference's bincode layout of ANNIndex (lsh.rs:47-55), from the
        device state (valid after add() as well)"""
        from . import bincode

        values, ids = self.values_and_ids()
        bincode.write_ann(file_path, self.max_node_size, [self.flatten(t) for t in range(self.info()["num_trees"])],
                          values, ids)

    @classmethod
    def load_index(cls, file_path: str, dim: int, *, seed: int = 4, ctx: Optional[Context] = None) -> "ANNIndex":
        """Index::load_index (base.rs:45-58) + the device forest (vers_lsh_from_parts).  N is a const generic in the
        reference, not stored in the file: the caller supplies ``dim``.  ``seed`` only feeds the sample pairs of leaf
        splits caused by later add() calls (the reference draws them from thread_rng)."""
        from . import bincode

        ctx = ctx or default_context()
        max_node_size, trees, values, ids = bincode.read_ann(file_path, dim)
        kind, leaf_len, planes, consts, items, tree_nodes = [], [], [], [], [], []
        for root in trees:  # preorder (node, ABOVE = right_node, BELOW = left_node), like vers_lsh_flatten
            n0 = len(kind)
            stack = [root]
            while stack:
                node = stack.pop()
                if node[0] == "leaf":
                    kind.append(1)
                    leaf_len.append(node[1].shape[0])
                    items.append(node[1])
                else:
                    kind.append(0)
                    leaf_len.append(0)
                    planes.append(node[1])
                    consts.append(node[2])
                    stack.append(node[3])  # left (below): popped second
                    stack.append(node[4])  # right (above): popped first
            tree_nodes.append(len(kind) - n0)
        kind = np.asarray(kind, np.uint8)
        leaf_len = np.asarray(leaf_len, np.uint32)
        planes = np.ascontiguousarray(np.asarray(planes, np.float32).reshape(-1, dim))
        consts = np.asarray(consts, np.float32)
        items = np.concatenate(items).astype(np.uint32) if items else np.empty(0, np.uint32)
        tree_nodes = np.asarray(tree_nodes, np.uint32)
        values = np.ascontiguousarray(values, np.float32)
        ids = np.ascontiguousarray(ids, np.uint64)
        h = C.c_void_p()
        check(lib().vers_lsh_from_parts(ctx.h, ptr(values), values.shape[0], dim, dim, ptr(ids), len(trees),
                                        max_node_size, seed, ptr(tree_nodes), ptr(kind), ptr(leaf_len), ptr(planes),
                                        ptr(consts), ptr(items), C.byref(h)))
        self = cls(ctx, h, dim)
        self.max_node_size = max_node_size
        return self

    def add(self, embedding, vec_id: int):
        """Index::add (lsh.rs:255-263)"""
        e = np.ascontiguousarray(embedding, np.float32).reshape(-1)
        check(lib().vers_lsh_add(self.h, ptr(e), vec_id))

    def search_approximate(self, query, top_k: int) -> List[Tuple[int, float]]:
        """Index::search_approximate (lsh.rs:264-282)"""
        ids, d, cnt = self.search_batch(query, top_k)
        return _to_pairs(ids[0], d[0], int(cnt[0]))

    def search_batch(self, queries, top_k: int):
        q = _rows(queries)
        ids, d, cnt = _result_buffers(q.shape[0], top_k)
        check(lib().vers_lsh_search(self.h, ptr(q), q.shape[0], q.shape[1], top_k, ptr(ids), ptr(d), ptr(cnt)))
        return ids[:, :top_k], d[:, :top_k], cnt
