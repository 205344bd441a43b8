"""diagnostic: latency of Index::add and of the one-query calls on a mid-size index"""
import sys, os, time, numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import vers_b200 as vb
ctx = vb.Context(0)
n, dim, C = 2_000_000, 768, 4096
ds = vb.Dataset.synth(ctx, 1, n, dim, kind=1, n_centers=16384, center_seed=7, row0=0, normalize=True)
init = vb.synth_init_rows(3, 1, C, n)
idx = vb.IVFFlatIndex.build_index(C, 1, 2, None, init_rows=init, ctx=ctx, dataset=ds)
extra = vb.Dataset.synth(ctx, 5, 600, dim, kind=1, n_centers=16384, center_seed=7, row0=0, normalize=True).download()
for i in range(50):
    idx.add(extra[i], 0)
t0 = time.perf_counter()
for i in range(50, 550):
    idx.add(extra[i], 0)
print(f"Index::add: {(time.perf_counter() - t0) / 500 * 1e6:.1f} us per call")
