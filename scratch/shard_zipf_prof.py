"""Profiling driver: the range-sharded join on Zipf-skewed probe keys with both ranks on ONE GPU (in-process group), so that ncu
can capture the heavy-hitter variants of the shuffle / partition kernels on a 1-GPU box.
    python scratch/shard_zipf_prof.py [log2_build] [log2_probe] [chunks] [reps] [zipf_s]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_olap_engine_b200 as g  # noqa: E402
from gpu_olap_engine_b200.shard import ShardGroup  # noqa: E402

lb = int(sys.argv[1]) if len(sys.argv) > 1 else 26
lp = int(sys.argv[2]) if len(sys.argv) > 2 else 28
chunks = int(sys.argv[3]) if len(sys.argv) > 3 else 2
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
zipf = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
world = 2
nb, npb = 1 << lb, 1 << lp
with ShardGroup([0] * world, nb, npb, chunks) as grp:
    for c in grp.ctxs:
        c.set_profiling(True)
    db = [grp.ctxs[r].generate(nb, 1, 3, nb * world, row0=r * nb) for r in range(world)]
    dp = [grp.ctxs[r].generate_fk(npb, 3, nb * world, 33, zipf, row0=r * npb) for r in range(world)]
    mm = [int(npb * 2.5) + 4096] * world
    out = [grp.ctxs[r].alloc(2 * mm[r], np.uint32) for r in range(world)]
    hot = [grp.ctxs[r].alloc(2 * npb, np.uint32) for r in range(world)] if zipf > 0 else None
    for i in range(reps):
        t0 = time.perf_counter()
        res = grp.join(db, [nb] * world, dp, [npb] * world, out, mm, hot_pairs=hot, max_hot_pairs=[npb] * world if hot else None)
        dt = time.perf_counter() - t0
        counts, hot_counts = res if hot else (res, [0] * world)
        assert sum(counts) + sum(hot_counts) == npb * world, (counts, hot_counts)
        print(f"rep {i}: {dt * 1e3:.2f} ms wall, pairs {counts}, hot pairs {hot_counts}, phases rank 0: "
              + ", ".join(f"{k} {v:.2f}" for k, v in grp.ctxs[0].last_phase_times().items()), flush=True)
