"""diagnostic: latency of the hyperplane forest's one-query search and of add"""
import sys, os, time, numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import vers_b200 as vb
ctx = vb.Context(0)
n, dim = 1_000_000, 300
rows = vb.Dataset.synth(ctx, 1, n, dim, kind=1, n_centers=65536, center_seed=7, row0=0, normalize=True).download()
idx = vb.ANNIndex.build_index(16, 100, rows, None, seed=4, ctx=ctx)
q = vb.Dataset.synth(ctx, 2, 300, dim, kind=1, n_centers=65536, center_seed=7, row0=0, normalize=True).download()
for i in range(20):
    idx.search_approximate(q[i], 10)
t0 = time.perf_counter()
for i in range(200):
    idx.search_batch(q[i:i + 1], 10)
print(f"ANNIndex::search_approximate (one query): {(time.perf_counter() - t0) / 200 * 1e6:.1f} us per call")
extra = vb.Dataset.synth(ctx, 5, 300, dim, kind=1, n_centers=65536, center_seed=7, row0=0, normalize=True).download()
for i in range(20):
    idx.add(extra[i], n + i)
t0 = time.perf_counter()
for i in range(20, 220):
    idx.add(extra[i], n + i)
print(f"ANNIndex::add: {(time.perf_counter() - t0) / 200 * 1e6:.1f} us per call")
