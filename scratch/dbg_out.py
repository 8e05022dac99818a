import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import gpu_olap_engine_b200 as g
ctx = g.Context(0); ctx.set_profiling(True)
nb, npb = 1 << 27, 1 << 29
db = ctx.generate(nb, 1, 3, nb); dp = ctx.generate(npb, 0, 33, nb)
pairs = ctx.alloc(2 * npb, np.uint32)
# probe keys that never match (domain shifted)
dmiss = ctx.generate(npb, 0, 34, nb); 
miss = ctx.project_i64(dmiss, '+', npb, literal=1 << 40)
for name, probe, want in [("pairs", dp, True), ("count-only", dp, False), ("no-match pairs", miss, True)]:
    for i in range(3):
        c, _ = ctx.hash_join_i64(db, nb, probe, npb, max_matches=npb, strategy=g.JOIN_RADIX, out_pairs=pairs if want else None, want_pairs=want)
    print(name, c, {k: round(v, 2) for k, v in ctx.last_phase_times().items()})
