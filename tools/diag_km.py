import sys, os, numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import oracle as vo, vers_b200 as vb
ctx = vb.Context(0)
n, ncent, C, dim = 4_000_000, 5242, 16384, 128
ds = vb.Dataset.synth(ctx, 1, n, dim, kind=1, n_centers=ncent, center_seed=7, row0=0, normalize=False)
init = vb.synth_init_rows(3, 1, C, n)[0]
m = 50_000
ds2 = vb.Dataset.synth(ctx, 1, m, dim, kind=1, n_centers=ncent, center_seed=7, row0=0, normalize=False)
ex = None
for mode in (3, 0, 2):
    km = vb.KMeans(ds, C); km.set_mode(mode); km.init_from_rows(init); km.assign_step()
    fl, a = km.last_uncertified_rows, km.assignments(); cents = km.centroids(); km.close()
    if ex is None:
        km2 = vb.KMeans(ds2, C); km2.set_mode(1); km2.set_centroids(cents); km2.assign_step()
        ex = km2.assignments(); km2.close()
    bad = np.flatnonzero(a[:m] != ex)
    print(f"mode {mode}: flagged {fl}, mismatch vs exact on {m}: {len(bad)}, first bad rows {bad[:6]}", flush=True)
