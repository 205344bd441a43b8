"""Generates tests/golden/c1_small.npz from the CPU oracle (python tests/golden/make_golden.py).

The reference ships no golden vectors and cannot be run here (nightly Rust, no toolchain), so these vectors pin the
ORACLE against accidental change (compiler flags, refactors); the oracle itself is pinned against an independent
numpy restatement in tests/test_oracle.py.  BASELINE config 1 shape, scaled to a few hundred KB."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import oracle as vo  # noqa: E402

bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)
P = dict(seed_data=1, seed_query=2, seed_init=3, seed_lsh=4, seed_centers=7, n=4000, dim=300, nq=24, C=16, k=10,
         max_iter=20, n_centers=16, trees=4, max_size=100)
rows = vo.synth(P["seed_data"], P["n"], P["dim"], kind=1, n_centers=P["n_centers"], center_seed=P["seed_centers"])
q = vo.synth(P["seed_query"], P["nq"], P["dim"], kind=1, n_centers=P["n_centers"], center_seed=P["seed_centers"])
init = vo.init_rows(P["seed_init"], 1, P["C"], P["n"])
cents, assign, cost, _ = vo.ivf_build_index(rows, P["C"], 1, P["max_iter"], init)
off, lr = vo.ivf_lists(assign, P["C"])
out = dict(P)
out.update(rows_head_bits=bits(rows[:8]), init_rows=init[0], assign=assign, cents_bits=bits(cents),
           cost_bits=bits(cost))
for name, nprobe in (("ref", 0), ("np4", 4)):
    ids, d, _ = vo.ivf_search(rows, cents, off, lr, q, P["k"], nprobe=nprobe)
    out[f"ids_{name}"], out[f"d_{name}_bits"] = ids, bits(d)
ids, d, _ = vo.exhaustive(rows, q, P["k"])
out["ids_flat"], out["d_flat_bits"] = ids, bits(d)
rng = np.random.default_rng(0)
planes = np.empty((16, P["dim"]), np.float32)
consts = np.empty(16, np.float32)
for p in range(16):
    a, b = rng.choice(P["n"], 2, replace=False)
    planes[p], consts[p] = vo.lsh_make_plane(rows[a], rows[b])
out["planes_bits"], out["consts_bits"] = bits(planes), bits(consts)
out["hash_packed"] = np.packbits(vo.lsh_hash(rows, planes, consts))
L = vo.LSH(rows, None, P["trees"], P["max_size"], P["seed_lsh"])
ids, d, _ = L.search(q, P["k"])
out["ids_lsh"], out["d_lsh_bits"] = ids, bits(d)
np.savez_compressed(os.path.join(os.path.dirname(__file__), "c1_small.npz"), **out)
print("wrote c1_small.npz")
