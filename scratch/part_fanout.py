# partition time vs fan-out (run length per partition per tile): is the scatter DRAM-locality bound?
import sys, numpy as np, torch
sys.path.insert(0, ".")
from gpu_olap_engine_b200.core import Context
ctx = Context()
n = 1 << 29
keys = ctx.generate(n, 0, 7, domain=1 << 27)
st = torch.cuda.ExternalStream(ctx.lib.olap_ctx_stream(ctx.handle, 0))
ok = ctx.alloc(n, np.int64); oi = ctx.alloc(n, np.uint32); begin = ctx.alloc(257, np.uint32)
import gpu_olap_engine_b200._lib as L
for bits in (2, 4, 5, 6, 7, 8):
    ts = []
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        L.check(ctx.lib.olap_radix_partition_i64(ctx.handle, keys.ptr, None, n, bits, ok.ptr, oi.ptr, begin.ptr, None))
        e1.record(st); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("bits", bits, "hist+scan+partition ms", [round(t, 3) for t in ts], flush=True)
