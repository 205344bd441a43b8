import numpy as np, sys
sys.path.insert(0,'/root/repo')
import oracle as vo, vers_b200 as vb
ctx = vb.Context(0)
n, dim, C = 10000, 300, 16
rows = vo.synth(1, n, dim, kind=1, n_centers=16, center_seed=7)
init = vo.init_rows(3, 1, C, n)
idx = vb.IVFFlatIndex.build_index(C, 1, 20, rows, init_rows=init, ctx=ctx)
q = vo.synth(2, 100, dim, kind=1, n_centers=16, center_seed=7)
cents, assign, _, _ = vo.ivf_build_index(rows, C, 1, 20, init)
off, lr = vo.ivf_lists(assign, C)
for mode in (0, 3, 2, 1):
    idx.set_mode(mode)
    ids, d, cnt = idx.search_batch(q, 10, nprobe=4)
    oi, od, oc = vo.ivf_search(rows, cents, off, lr, q, 10, nprobe=4)
    print("mode", mode, "ids equal", np.array_equal(ids, oi), idx.last_search_stats())
