#!/usr/bin/env python
"""Secondary benchmarks (the other BASELINE configs): group-by cardinality sweep (C3), fused filter (C1/C4 shapes), and the
TPC-H Q6 / Q1 shapes (C4) through the operator API (olap_execute on device-resident tables).
Prints one JSON line per measurement; bench.py stays the contract benchmark.   python bench_ops.py [--log2-rows 30] [--sf 10]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2-rows", type=int, default=30)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--groups", default="1,8,100,1000,10000,100000,1000000,10000000,100000000")
    ap.add_argument("--only", default="groupby,filter,tpch")
    ap.add_argument("--sf", type=float, default=10.0, help="TPC-H scale factor of the lineitem-shaped table (6,000,379 rows per unit)")
    args = ap.parse_args()
    import numpy as np
    import torch
    import gpu_olap_engine_b200 as g
    from bench import peak_hbm
    peak, kind = peak_hbm()
    ctx = g.Context(0, 0, 8)
    ctx.set_profiling(True)
    stream = torch.cuda.ExternalStream(ctx.lib.olap_ctx_stream(ctx.handle, 0))
    n = 1 << args.log2_rows

    def timed(fn):
        for _ in range(args.warmup):
            r = fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(stream)
        for _ in range(args.steps):
            r = fn()
        e1.record(stream); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps, r

    vals = ctx.generate(n, 0, 5, 1 << 20)          # i64 uniform [0, 2^20): SUM < 2^50 stays exact
    if "groupby" in args.only:
        keys = ctx.alloc(n, np.int64)
        for G in [int(x) for x in args.groups.split(",")]:
            ctx.generate(n, 0, 4, G, out=keys)
            aggs = [(g.SUM, vals, g.I64), (g.COUNT, vals, g.I64), (g.MIN, vals, g.I64), (g.MAX, vals, g.I64)]
            cap = min(G, n)
            ok = ctx.alloc(cap, np.int64); oa = ctx.alloc(4 * cap, np.uint64)
            import ctypes as C
            from gpu_olap_engine_b200._lib import AggSpec, check
            specs = (AggSpec * 4)(*[AggSpec(v.ptr, f, t) for f, v, t in aggs])

            def run():
                gg = C.c_uint64()
                check(ctx.lib.olap_group_by_i64(ctx.handle, keys.ptr if G > 1 else None, specs, 4, n, G, None, 0, ok.ptr, oa.ptr, cap, None, C.byref(gg), None))
                return gg.value
            ms, ng = timed(run)
            by = 16.0 * n if G > 1 else 8.0 * n
            print(json.dumps({"op": "group_by SUM/COUNT/MIN/MAX i64", "rows": n, "groups_hint": G, "groups_found": int(ng), "ms": ms,
                              "rows_per_s": n / ms * 1e3, "algorithmic_gbs": by / ms / 1e6, "frac_of_peak": by / ms / 1e6 / peak,
                              "peak": peak, "peak_kind": kind, "phases_ms": ctx.last_phase_times()}), flush=True)
            del ok, oa
        del keys
    if "filter" in args.only:
        col = vals
        mask = ctx.alloc(n, np.uint8); sel = ctx.alloc(n, np.uint32)
        import ctypes as C
        from gpu_olap_engine_b200._lib import check
        for name, m, s in [("mask only", mask, None), ("mask + selection vector", mask, sel), ("selection vector only", None, sel)]:
            def run():
                c = C.c_uint64()
                check(ctx.lib.olap_filter_i64(ctx.handle, col.ptr, n, g.GT, 1 << 19, m.ptr if m else None, s.ptr if s else None, None, C.byref(c), None))
                return c.value
            ms, kept = timed(run)
            by = 8.0 * n + (n if m else 0) + (4.0 * kept if s else 0)
            print(json.dumps({"op": f"filter i64 > literal ({name})", "rows": n, "kept": int(kept), "ms": ms, "rows_per_s": n / ms * 1e3,
                              "algorithmic_gbs": by / ms / 1e6, "frac_of_peak": by / ms / 1e6 / peak}), flush=True)
    if "tpch" in args.only:
        del vals
        bench_tpch(ctx, g, args, peak)


def bench_tpch(ctx, g, args, peak):
    """C4: lineitem-shaped columns, registered once as device-resident batches; Q6 and Q1 as PhysicalPlan
    trees through olap_execute.  Time = wall clock around execute() (plan lowering, kernels, export of the tiny result)."""
    import time
    import numpy as np
    import pyarrow as pa
    n = int(6_000_379 * args.sf)
    rng = np.random.default_rng(7)
    qty = rng.integers(1, 51, n).astype(np.float64)
    price = np.round(qty * (900.0 + rng.integers(0, 120001, n) / 100.0) * 100.0) / 100.0
    cols = {"l_shipdate": rng.integers(8036, 10562, n, dtype=np.int64), "l_quantity": qty, "l_extendedprice": price,
            "l_discount": rng.integers(0, 11, n) / 100.0, "l_tax": rng.integers(0, 9, n) / 100.0,
            "l_returnflag": rng.integers(0, 3, n, dtype=np.int64), "l_linestatus": rng.integers(0, 2, n, dtype=np.int64)}
    ex = g.GpuExecutor(g.EngineConfig(), ctx=ctx)
    ex.drop_table("lineitem")
    step = 1 << 24
    for a in range(0, n, step):
        ex.register_table("lineitem", pa.record_batch({k: v[a:a + step] for k, v in cols.items()}))
    del cols, qty, price
    col = lambda name: g.Column(0, name)
    lit = lambda v: g.Literal(str(v))
    B = g.BinaryExpr
    AND = lambda a, b: B(a, "AND", b)
    q6 = g.GpuAggregate(g.GpuFilter(g.GpuTableScan("lineitem"),
                                    AND(AND(B(col("l_shipdate"), ">=", lit(8766)), B(col("l_shipdate"), "<", lit(9131))),
                                        AND(AND(B(col("l_discount"), ">=", lit("0.05")), B(col("l_discount"), "<=", lit("0.07"))),
                                            B(col("l_quantity"), "<", lit("24.0"))))),
                        [], [g.Sum(B(col("l_extendedprice"), "*", col("l_discount")))])
    one = lit("1.0")
    disc_price = B(col("l_extendedprice"), "*", B(one, "-", col("l_discount")))
    charge = B(disc_price, "*", B(one, "+", col("l_tax")))
    key = B(B(col("l_returnflag"), "*", lit(2)), "+", col("l_linestatus"))
    q1 = g.GpuAggregate(g.GpuFilter(g.GpuTableScan("lineitem"), B(col("l_shipdate"), "<=", lit(10471))), [key],
                        [g.Sum(col("l_quantity")), g.Sum(col("l_extendedprice")), g.Sum(disc_price), g.Sum(charge), g.Avg(col("l_quantity")),
                         g.Avg(col("l_extendedprice")), g.Avg(col("l_discount")), g.Count(col("l_quantity"))])
    for name, plan, ncols in (("TPC-H Q6 shape", q6, 4), ("TPC-H Q1 shape", q1, 7)):
        for _ in range(args.warmup):
            out = ex.execute(plan)
        ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out = ex.execute(plan)
        ctx.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / args.steps
        by = 8.0 * ncols * n
        print(json.dumps({"op": f"{name} via olap_execute (GpuAggregate(GpuFilter(GpuTableScan)))", "rows": n, "sf": args.sf, "result_rows": out[0].num_rows,
                          "ms": ms, "rows_per_s": n / ms * 1e3, "algorithmic_gbs": by / ms / 1e6, "frac_of_peak": by / ms / 1e6 / peak,
                          "launches": None}), flush=True)
    ex.drop_table("lineitem")


if __name__ == "__main__":
    main()
