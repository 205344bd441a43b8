"""CPU tier: the C-ABI library loads, exports every symbol include/vers_device.h declares, and fails loudly
(no CPU fallback) when there is no CUDA device.  No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "vers_device.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vers_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(vb):
    syms = header_symbols()
    assert len(syms) >= 50
    out = subprocess.run(["nm", "-D", "--defined-only", vb.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (vers_[a-z0-9_]+)", out))
    missing = [s for s in syms if s not in exported]
    assert not missing, f"declared but not exported: {missing}"
    # and the ctypes table covers the header exactly
    assert sorted(vb._abi.SIGNATURES) == syms


def test_library_is_sm100a_and_independent_of_the_oracle(vb):
    out = subprocess.run(["cuobjdump", "--list-elf", vb.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "sm_90" not in out and "sm_80" not in out
    ldd = subprocess.run(["ldd", vb.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "torch" not in ldd


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "vers_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "vers_oracle" not in src, f


def test_abi_version_and_error_string(vb):
    assert vb.lib().vers_abi_version() == 1
    assert isinstance(vb.lib().vers_last_error(), (bytes, type(None)))


def test_fails_loudly_without_a_gpu(vb):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(vb.VersError) as e:
        vb.Context(0)
    assert e.value.code == vb.ERR_CUDA and "no CPU fallback" in str(e.value)


def test_null_arguments_are_rejected_not_crashed(vb):
    L = vb.lib()
    assert L.vers_ctx_create(0, None) == vb.ERR_ARG
    assert L.vers_dataset_normalize(None) == vb.ERR_ARG
    assert L.vers_ivf_search(None, None, 1, 1, 1, 0, None, None, None) == vb.ERR_ARG
    assert L.vers_kmeans_fit(None, 1, None) == vb.ERR_ARG
    assert L.vers_lsh_hash(None, None, 1, 1, None, None) == vb.ERR_ARG
    assert b"null" in L.vers_last_error()


def test_bincode_layout_roundtrip(tmp_path):
    """Index::save_index layout (base.rs:31-43; struct at ivfflat.rs:9-15) without touching the GPU"""
    from vers_b200.bincode import read_ivfflat, write_ivfflat

    rng = np.random.default_rng(0)
    values = rng.standard_normal((50, 7)).astype(np.float32)
    cents = rng.standard_normal((4, 7)).astype(np.float32)
    assign = rng.integers(0, 4, 50).astype(np.uint64)
    assign[assign == 2] = 1  # an empty list
    p = str(tmp_path / "x.bin")
    write_ivfflat(p, 4, values, cents, assign)
    raw = open(p, "rb").read()
    # hand-check the head of the layout: u64 num_centroids, u64 len(values), first row as 7 raw f32
    assert int.from_bytes(raw[0:8], "little") == 4 and int.from_bytes(raw[8:16], "little") == 50
    assert np.array_equal(np.frombuffer(raw, "<f4", 7, 16), values[0])
    expected = 8 + (8 + 50 * 28) + (8 + 4 * 28) + (8 + 50 * 8) + 8 + 4 * 8 + 50 * 8
    assert len(raw) == expected
    nc, v2, c2, a2, ids = read_ivfflat(p)
    assert nc == 4 and np.array_equal(v2, values) and np.array_equal(c2, cents) and np.array_equal(a2, assign)
    assert len(ids[2]) == 0
    for c in range(4):
        assert np.array_equal(ids[c], np.nonzero(assign == c)[0].astype(np.uint64))


def test_cpp_host_mirror_compiles_and_links(vb, tmp_path):
    """the C++ host mirror builds against the header and links against the library (no GPU needed)"""
    libdir = os.path.dirname(vb.LIB_PATH)
    exe = str(tmp_path / "host_check")
    subprocess.run(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "host", "host_check.cpp"), "-o", exe, "-L" + libdir,
                    "-lvers_b200", "-Wl,-rpath," + libdir], check=True)
    assert os.path.exists(exe)


def test_bincode_ann_index_layout(tmp_path, vo):
    """ANNIndex (lsh.rs:13-55) in the reference's bincode layout: written from the flattened forest (the oracle's here,
    the device forest is compared with it bit for bit in the GPU tests), read back, and checked node by node against
    an independent walk of the flattened preorder (node, above, below) -> serialized order (node, left=below,
    right=above)."""
    import struct

    from vers_b200.bincode import read_ann, write_ann

    n, dim, T, max_size = 700, 12, 3, 9
    rows = vo.synth(1, n, dim, kind=1, n_centers=8, center_seed=7)
    rows[50] = rows[3]  # dropped by deduplicate (lsh.rs:113-130)
    ids = np.arange(n, dtype=np.uint64) + 100
    forest = vo.LSH(rows, ids, T, max_size, 4)
    keep = np.array([i for i in range(n) if i != 50])
    assert forest.num_values == keep.shape[0]
    flats = [forest.flatten(t) for t in range(T)]
    path = str(tmp_path / "ann.bin")
    write_ann(path, max_size, flats, rows[keep], ids[keep])
    mns, trees, values, rid = read_ann(path, dim)
    assert mns == max_size and len(trees) == T
    assert np.array_equal(values.view(np.uint32), rows[keep].view(np.uint32)) and np.array_equal(rid, ids[keep])

    def walk(flat):
        pos = {"node": 0, "plane": 0, "item": 0}

        def rec():
            i = pos["node"]
            pos["node"] += 1
            if flat["kind"][i] == 1:
                ln = int(flat["leaf_len"][i])
                it = flat["items"][pos["item"]:pos["item"] + ln].astype(np.uint64)
                pos["item"] += ln
                return ("leaf", it)
            p = pos["plane"]
            pos["plane"] += 1
            above = rec()
            below = rec()
            return ("inner", flat["planes"][p], flat["consts"][p], below, above)  # left = below, right = above

        return rec()

    def same(a, b):
        if a[0] != b[0]:
            return False
        if a[0] == "leaf":
            return np.array_equal(a[1], b[1])
        return (np.array_equal(np.asarray(a[1], np.float32).view(np.uint32), np.asarray(b[1], np.float32).view(np.uint32))
                and np.float32(a[2]).view(np.uint32) == np.float32(b[2]).view(np.uint32) and same(a[3], b[3])
                and same(a[4], b[4]))

    for t in range(T):
        assert same(trees[t], walk(flats[t])), f"tree {t}"
    # header bytes: max_node_size, number of trees, then the first Node's variant tag (u32)
    raw = open(path, "rb").read()
    assert struct.unpack_from("<QQI", raw, 0) == (max_size, T, int(flats[0]["kind"][0]))
    # every leaf holds fewer than max_size rows (lsh.rs:97-98) and every row appears exactly once per tree
    def leaves(nd):
        return [nd[1]] if nd[0] == "leaf" else leaves(nd[3]) + leaves(nd[4])

    for t in range(T):
        ls = leaves(trees[t])
        assert all(l.shape[0] < max_size for l in ls)
        assert np.array_equal(np.sort(np.concatenate(ls)), np.arange(keep.shape[0], dtype=np.uint64))


def test_rust_sys_crate_declares_the_whole_header():
    """rust/vers-cuda-sys/src/lib.rs is generated from include/vers_device.h (tools/gen_rust_sys.py): it must be up to
    date and declare every entry point the header (and the ctypes table) does"""
    import re
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert subprocess.run([sys.executable, os.path.join(root, "tools", "gen_rust_sys.py"), "--check"]).returncode == 0, \
        "rust/vers-cuda-sys/src/lib.rs is stale: run tools/gen_rust_sys.py"
    rs = open(os.path.join(root, "rust", "vers-cuda-sys", "src", "lib.rs")).read()
    declared = set(re.findall(r"pub fn (vers_\w+)\(", rs))
    from vers_b200 import _abi

    assert declared == set(_abi.SIGNATURES), sorted(declared ^ set(_abi.SIGNATURES))
    used = set(re.findall(r"sys::(vers_\w+)\(", open(os.path.join(root, "rust", "vers-gpu", "src", "lib.rs")).read()))
    assert used <= declared, sorted(used - declared)
