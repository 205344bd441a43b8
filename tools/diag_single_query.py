"""diagnostic: kernel composition of the reference's one-query call (run under ncu --metrics gpu__time_duration.sum)"""
import sys, os, numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import vers_b200 as vb
ctx = vb.Context(0)
n, dim, C = 10_000_000, 768, 4096
ds = vb.Dataset.synth(ctx, 1, n, dim, kind=1, n_centers=65536, center_seed=7, row0=0, normalize=True)
init = vb.synth_init_rows(3, 1, C, n)
idx = vb.IVFFlatIndex.build_index(C, 1, 2, None, init_rows=init, ctx=ctx, dataset=ds)
q = vb.Dataset.synth(ctx, 2, 8, dim, kind=1, n_centers=65536, center_seed=7, row0=0, normalize=True).download()
for i in range(4):
    idx.search_batch(q[i:i + 1], 10, nprobe=0)
print("MARK")
for i in range(4):
    idx.search_batch(q[i:i + 1], 10, nprobe=32)
