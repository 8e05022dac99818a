# C2-shaped join with (a) dense build keys 0..n-1 (the benchmark's generator) and (b) a sparse random subset of a 64x larger domain
import sys, numpy as np, torch
sys.path.insert(0, ".")
from gpu_olap_engine_b200.core import Context
import gpu_olap_engine_b200._lib as L
ctx = Context()
nb, npb = 1 << 27, 1 << 29
st = torch.cuda.ExternalStream(ctx.lib.olap_ctx_stream(ctx.handle, 0))
class Cai:
    def __init__(s, ptr, n, t): s.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": (n,), "typestr": t, "version": 3}
for name, dom in (("dense", nb), ("sparse x64", nb * 64)):
    db = ctx.generate(nb, 1, 3, dom)
    idx = ctx.generate(npb, 0, 33, nb)
    dp = ctx.alloc(npb, np.int64)
    ctx.synchronize()
    b = torch.as_tensor(Cai(db.ptr, nb, "<i8"), device="cuda"); i = torch.as_tensor(Cai(idx.ptr, npb, "<i8"), device="cuda")
    p = torch.as_tensor(Cai(dp.ptr, npb, "<i8"), device="cuda")
    torch.index_select(b, 0, i, out=p); torch.cuda.synchronize()
    del idx, i
    pairs = ctx.alloc(2 * npb, np.uint32)
    ts = []
    for rep in range(5):
        ctx.set_profiling(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        cnt, _ = ctx.hash_join_i64(db, nb, dp, npb, max_matches=npb, strategy=0, out_pairs=pairs)
        e1.record(st); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1)); ph = ctx.last_phase_times()
    assert cnt == npb
    print(name, "ms", [round(t, 2) for t in ts], {k: round(v, 2) for k, v in ph.items()}, flush=True)
    del db, dp, pairs, b, p
