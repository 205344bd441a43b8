"""Generates tests/golden/simd_small.npz from the CPU oracle (python tests/golden/make_golden_simd.py): HNSW's SIMD-order
distances (base.rs:158-294) for a few hundred (query, row) pairs at the reference's dimension (300 = 4 x 64 + 11 x 4) and
at one with a scalar tail (67 = 64 + 0 x 4 + 3).  Pins the oracle against accidental change; the oracle itself is pinned
against the numpy restatement in tests/test_oracle.py."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import oracle as vo  # noqa: E402

bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)
out = {}
for dim in (300, 67):
    rows = vo.synth(1, 500, dim, kind=1, n_centers=8, center_seed=7, normalize=True)
    q = vo.synth(2, 6, dim, kind=1, n_centers=8, center_seed=7, normalize=True)
    rng = np.random.default_rng(dim)
    pr = rng.integers(0, 500, 300).astype(np.uint64)
    pq = rng.integers(0, 6, 300).astype(np.uint32)
    out[f"pair_row_{dim}"], out[f"pair_query_{dim}"] = pr, pq
    for metric in (0, 1):
        out[f"d_{dim}_m{metric}_bits"] = bits(vo.pair_distances_simd(rows, q, pr, pq, metric))
np.savez_compressed(os.path.join(os.path.dirname(__file__), "simd_small.npz"), **out)
print("wrote simd_small.npz")
