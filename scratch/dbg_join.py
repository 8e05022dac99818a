import sys, numpy as np
sys.path.insert(0, '.')
import gpu_olap_engine_b200 as g, oracle as O
ctx = g.Context(0)
for (nb, npb, dom, strat) in [(100000, 1000000, 100000, 2), (100000, 1000000, 100000, 0), (100000, 200000, 100000, 2), (100000, 1000000, 10**9, 2)]:
    b = O.gen_uniform_i64(nb, 11, dom); p = O.gen_uniform_i64(npb, 12, dom)
    ec, ep = O.hash_join_i64(b, p)
    db, dp = ctx.upload_raw(b), ctx.upload_raw(p)
    cnt, pairs = ctx.hash_join_i64(db, nb, dp, npb, max_matches=max(ec, 1), strategy=strat)
    got = pairs.to_numpy(2 * min(cnt, max(ec, 1))).reshape(-1, 2)
    print(nb, npb, dom, strat, "count", cnt, "expected", ec)
    if cnt != ec:
        gp = np.bincount(got[:, 1], minlength=npb); xp = np.bincount(ep[:, 1], minlength=npb)
        bad = np.nonzero(gp != xp)[0]
        print(" probe rows with wrong match count:", len(bad), bad[:10], gp[bad[:10]], xp[bad[:10]])
        print(" bad rows mod 1024:", np.unique(bad % 1024)[:20], " tiles:", np.unique(bad // 1024)[:20])
        ok = (b[got[:, 0]] == p[got[:, 1]]).all(); print(" all emitted pairs valid:", ok)
