"""Profiling driver: the range-sharded join with both ranks on ONE GPU (in-process group), so that ncu can capture the shuffle
kernel, the second partition level and the shared-memory join of the multi-GPU path on a 1-GPU box.
    python scratch/shard_prof.py [log2_build] [log2_probe] [chunks] [reps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_olap_engine_b200 as g  # noqa: E402
from gpu_olap_engine_b200.shard import ShardGroup  # noqa: E402

lb = int(sys.argv[1]) if len(sys.argv) > 1 else 25
lp = int(sys.argv[2]) if len(sys.argv) > 2 else 27
chunks = int(sys.argv[3]) if len(sys.argv) > 3 else 4
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
world = 2
nb, npb = 1 << lb, 1 << lp
with ShardGroup([0] * world, nb, npb, chunks) as grp:
    db = [grp.ctxs[r].generate(nb, 1, 3, nb * world, row0=r * nb) for r in range(world)]
    dp = [grp.ctxs[r].generate(npb, 0, 33, nb * world, row0=r * npb) for r in range(world)]
    mm = [int(npb * 1.3) + 4096] * world
    out = [grp.ctxs[r].alloc(2 * mm[r], np.uint32) for r in range(world)]
    for i in range(reps):
        t0 = time.perf_counter()
        counts = grp.join(db, [nb] * world, dp, [npb] * world, out, mm)
        dt = time.perf_counter() - t0
        assert sum(counts) == npb * world, counts
        print(f"rep {i}: {dt * 1e3:.2f} ms wall for {world} emulated ranks on one GPU, {(nb + npb) * world / dt / 1e9:.1f} G rows/s")
