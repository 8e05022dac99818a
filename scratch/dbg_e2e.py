import sys, time, numpy as np, torch, ctypes as C
sys.path.insert(0, '.')
import gpu_olap_engine_b200 as g
ctx = g.Context(0); ctx.set_profiling(True)
nb, npb = 1 << 27, 1 << 29
db = ctx.generate(nb, 1, 3, nb); dp = ctx.generate(npb, 0, 33, nb)
pairs = ctx.alloc(2 * npb, np.uint32)
stream = torch.cuda.ExternalStream(ctx.lib.olap_ctx_stream(ctx.handle, 0))
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3 / reps, r
for strat, name in [(g.JOIN_RADIX, "radix one-shot"), (g.JOIN_BROADCAST, "single global table one-shot")]:
    ms, (c, _) = timed(lambda: ctx.hash_join_i64(db, nb, dp, npb, max_matches=npb, strategy=strat, out_pairs=pairs))
    print(name, c, f"{ms:.2f} ms", {k: round(v, 2) for k, v in ctx.last_phase_times().items()})
for strat, name in [(g.JOIN_RADIX, "radix"), (g.JOIN_BROADCAST, "global table")]:
    for nbatch in (4, 8, 16):
        bsz = npb // nbatch
        def run():
            tot = 0
            with g.JoinBuild(ctx, db, nb, strat) as jb:
                for i in range(nbatch):
                    tot += jb.probe(dp.ptr + 8 * i * bsz, bsz, probe_row_base=i * bsz, out_pairs=pairs.ptr + 8 * tot, max_matches=npb - tot)
            return tot
        ms, c = timed(run, 2)
        print(f"batched {name} x{nbatch}: {c} {ms:.2f} ms")
