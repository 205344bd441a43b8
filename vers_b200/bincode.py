"""bincode 1.3.3 (default options: little endian, fixed-width ints, u64 lengths) codec for the reference's
``IVFFlatIndex<N>`` so an index saved by either side loads in the other (Index::save_index / load_index,
indexes/base.rs:31-58).

Field order as declared at indexes/ivfflat.rs:9-15:
    num_centroids: usize            -> u64
    values: Vec<Vector<N>>          -> u64 len, then len * N f32   (serde_arrays writes [f32; N] as an N-tuple: no
                                                                    length prefix, no padding: 4N bytes per row)
    centroids: Vec<Vector<N>>       -> same
    assignments: Vec<usize>         -> u64 len, then len * u64
    ids: Vec<Vec<usize>>            -> u64 len, then for each list u64 len + len * u64
N is a const generic in the reference and is NOT stored in the file; the reader recovers it from the sizes.
"""
from __future__ import annotations

import struct

import numpy as np


def write_ivfflat(path: str, num_centroids: int, values: np.ndarray, centroids: np.ndarray, assignments: np.ndarray):
    values = np.ascontiguousarray(values, "<f4")
    centroids = np.ascontiguousarray(centroids, "<f4")
    a = np.ascontiguousarray(assignments, "<u8")
    order = np.argsort(a, kind="stable")  # ids[c] = ascending row ids of cluster c (ivfflat.rs:123-127)
    bounds = np.searchsorted(a[order], np.arange(num_centroids + 1))
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", num_centroids))
        f.write(struct.pack("<Q", values.shape[0]))
        f.write(values.tobytes())
        f.write(struct.pack("<Q", centroids.shape[0]))
        f.write(centroids.tobytes())
        f.write(struct.pack("<Q", a.shape[0]))
        f.write(a.tobytes())
        f.write(struct.pack("<Q", num_centroids))
        for c in range(num_centroids):
            lst = order[bounds[c]:bounds[c + 1]].astype("<u8")
            f.write(struct.pack("<Q", lst.shape[0]))
            f.write(lst.tobytes())


def read_ivfflat(path: str):
    with open(path, "rb") as f:
        buf = f.read()
    total = len(buf)
    (num_centroids,) = struct.unpack_from("<Q", buf, 0)
    (n_values,) = struct.unpack_from("<Q", buf, 8)
    # N is not stored: solve for it.  After values (n*4N) comes a u64 that must equal num_centroids... we instead use
    # the trailing structure: find N such that the whole file parses exactly.
    def try_dim(N):
        off = 16 + n_values * 4 * N
        if off + 8 > total:
            return None
        (n_c,) = struct.unpack_from("<Q", buf, off)
        off += 8 + n_c * 4 * N
        if off + 8 > total:
            return None
        (n_a,) = struct.unpack_from("<Q", buf, off)
        if n_a != n_values:
            return None
        off += 8 + n_a * 8
        if off + 8 > total:
            return None
        (n_l,) = struct.unpack_from("<Q", buf, off)
        off += 8
        for _ in range(n_l):
            if off + 8 > total:
                return None
            (ln,) = struct.unpack_from("<Q", buf, off)
            off += 8 + ln * 8
        return n_c if off == total else None

    dim = None
    if n_values > 0:
        # upper bound on N from the file size
        for N in range(1, (total - 16) // (4 * n_values) + 1):
            if try_dim(N) is not None:
                dim = N
                break
    if dim is None:
        raise ValueError("not a bincode IVFFlatIndex file (cannot recover the vector dimension)")
    off = 16
    values = np.frombuffer(buf, "<f4", n_values * dim, off).reshape(n_values, dim).copy()
    off += n_values * 4 * dim
    (n_c,) = struct.unpack_from("<Q", buf, off)
    off += 8
    centroids = np.frombuffer(buf, "<f4", n_c * dim, off).reshape(n_c, dim).copy()
    off += n_c * 4 * dim
    (n_a,) = struct.unpack_from("<Q", buf, off)
    off += 8
    assignments = np.frombuffer(buf, "<u8", n_a, off).copy()
    off += n_a * 8
    (n_l,) = struct.unpack_from("<Q", buf, off)
    off += 8
    ids = []
    for _ in range(n_l):
        (ln,) = struct.unpack_from("<Q", buf, off)
        off += 8
        ids.append(np.frombuffer(buf, "<u8", ln, off).copy())
        off += ln * 8
    return num_centroids, values, centroids, assignments, ids
