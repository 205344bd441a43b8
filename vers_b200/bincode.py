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


# ---------------------------------------------------------------------------------------------- ANNIndex (lsh.rs:13-55)
#   max_node_size: usize                  -> u64
#   trees: Vec<Node<N>>                   -> u64 len, then every tree:
#       enum Node { Inner(Box<InnerNode>), Leaf(Box<LeafNode>) }  -> u32 variant (0 = Inner, 1 = Leaf), Box is transparent
#       InnerNode { hyperplane { coefficients: Vector<N> (N f32), constant: f32 }, left_node, right_node }
#                                         -> N f32, f32, then the LEFT (below) subtree, then the RIGHT (above) subtree
#       LeafNode(Vec<usize>)              -> u64 len, len * u64 (indices into `values`)
#   values: Vec<Vector<N>>                -> u64 len, len * N f32 (the deduplicated rows)
#   ids: Vec<usize>                       -> u64 len, len * u64
# A tree arrives here as the flattened preorder (node, ABOVE subtree, BELOW subtree) of vers_lsh_flatten / the oracle:
# kind[i] (0 inner, 1 leaf), leaf_len[i], planes / consts in preorder of the inner nodes, items concatenated in
# preorder of the leaves.


def _tree_to_bincode(t, dim: int) -> bytes:
    kind, leaf_len = t["kind"], t["leaf_len"]
    planes = np.ascontiguousarray(t["planes"], "<f4").reshape(-1, dim)
    consts = np.ascontiguousarray(t["consts"], "<f4")
    items = np.ascontiguousarray(t["items"], "<u8")
    nn = kind.shape[0]
    # pass 1 (iterative preorder walk): per node its plane index / item offset and the span of its subtree
    plane_of = np.full(nn, -1, np.int64)
    item_off = np.zeros(nn, np.int64)
    end = np.zeros(nn, np.int64)  # one past the last node of the subtree rooted at i
    pi = io = 0
    for i in range(nn):
        if kind[i] == 0:
            plane_of[i] = pi
            pi += 1
        else:
            item_off[i] = io
            io += int(leaf_len[i])
    stack = []  # (node, children still missing)
    for i in range(nn):
        if kind[i] == 0:
            stack.append([i, 2])
        else:
            end[i] = i + 1
            while stack:
                stack[-1][1] -= 1
                if stack[-1][1]:
                    break
                j, _ = stack.pop()
                end[j] = i + 1
    if stack:
        raise ValueError("truncated tree")
    # pass 2: emit (node, below, above); above subtree = [i+1, end[i+1]), below subtree = [end[i+1], end[i])
    out = []
    todo = [0]
    while todo:
        i = todo.pop()
        if kind[i] == 1:
            n = int(leaf_len[i])
            out.append(struct.pack("<IQ", 1, n))
            out.append(items[item_off[i]:item_off[i] + n].tobytes())
        else:
            out.append(struct.pack("<I", 0))
            out.append(planes[plane_of[i]].tobytes())
            out.append(consts[plane_of[i]:plane_of[i] + 1].tobytes())
            above, below = i + 1, int(end[i + 1])
            todo.append(above)   # popped second: right_node
            todo.append(below)   # popped first: left_node
    return b"".join(out)


def write_ann(path: str, max_node_size: int, trees, values: np.ndarray, ids: np.ndarray):
    """trees: list of flatten() dicts (one per tree)"""
    values = np.ascontiguousarray(values, "<f4")
    ids = np.ascontiguousarray(ids, "<u8")
    dim = values.shape[1]
    with open(path, "wb") as f:
        f.write(struct.pack("<QQ", max_node_size, len(trees)))
        for t in trees:
            f.write(_tree_to_bincode(t, dim))
        f.write(struct.pack("<Q", values.shape[0]))
        f.write(values.tobytes())
        f.write(struct.pack("<Q", ids.shape[0]))
        f.write(ids.tobytes())


def read_ann(path: str, dim: int):
    """-> (max_node_size, trees, values, ids); a tree is a nested tuple ("inner", coef, const, left, right) /
    ("leaf", items).  N is not stored in the file (const generic): the caller supplies it."""
    with open(path, "rb") as f:
        buf = f.read()
    off = 0

    def node():
        nonlocal off
        (v,) = struct.unpack_from("<I", buf, off)
        off += 4
        if v == 1:
            (n,) = struct.unpack_from("<Q", buf, off)
            off += 8
            it = np.frombuffer(buf, "<u8", n, off).copy()
            off += 8 * n
            return ("leaf", it)
        if v != 0:
            raise ValueError(f"bad Node variant {v}")
        coef = np.frombuffer(buf, "<f4", dim, off).copy()
        off += 4 * dim
        (c,) = struct.unpack_from("<f", buf, off)
        off += 4
        left = node()
        right = node()
        return ("inner", coef, np.float32(c), left, right)

    import sys

    sys.setrecursionlimit(max(sys.getrecursionlimit(), 100000))
    max_node_size, nt = struct.unpack_from("<QQ", buf, 0)
    off = 16
    trees = [node() for _ in range(nt)]
    (nv,) = struct.unpack_from("<Q", buf, off)
    off += 8
    values = np.frombuffer(buf, "<f4", nv * dim, off).reshape(nv, dim).copy()
    off += 4 * nv * dim
    (ni,) = struct.unpack_from("<Q", buf, off)
    off += 8
    ids = np.frombuffer(buf, "<u8", ni, off).copy()
    off += 8 * ni
    if off != len(buf):
        raise ValueError("trailing bytes: wrong dimension or not an ANNIndex file")
    return max_node_size, trees, values, ids
