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
    ap.add_argument("--sf", type=float, default=100.0, help="TPC-H scale factor of the lineitem-shaped table (6,000,379 rows per unit); BASELINE config 4 is SF100")
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


def device_lineitem(ctx, g, n):
    """lineitem-shaped columns of SURVEY.md 8(d) C4, generated ON THE DEVICE (olap_generate + the projection kernels): SF100 is 600 M
    rows x 7 columns = 33.6 GB, which a host-side generator + PCIe would turn into minutes of set-up.  {name: (DeviceBuffer, type)}."""
    import numpy as np
    from gpu_olap_engine_b200._lib import check
    lib, h = ctx.lib, ctx.handle

    def f64_op(col, op, lit, rhs=None):   # out-of-place (the projection kernels take __restrict__ pointers); the input is dropped
        out = ctx.alloc(n, np.float64)
        check(lib.olap_project_f64(h, col.ptr, rhs.ptr if rhs is not None else None, float(lit), op, n, out.ptr, None))
        ctx.synchronize()
        return out

    def f64_of_int(seed, domain):            # (u % domain) as f64: kind 2 gives (u % domain) / 1024 exactly, and x 1024 is exact too
        return f64_op(ctx.generate(n, 2, seed, domain), b"*", 1024.0)
    days = ctx.generate(n, 0, 70, 2526)
    shipdate = ctx.alloc(n, np.int64)
    check(lib.olap_project_i64(h, days.ptr, None, 8036, b"+", n, shipdate.ptr, None))                 # days 8036 .. 10561
    ctx.synchronize()
    del days
    qty = f64_op(f64_of_int(71, 50), b"+", 1.0)                                                       # 1 .. 50
    price = f64_op(f64_op(f64_op(f64_of_int(72, 120001), b"/", 100.0), b"+", 900.0), b"*", 0.0, rhs=qty)   # (900.00 .. 2100.00) x quantity
    discount = f64_op(f64_of_int(73, 11), b"/", 100.0)                                                # 0.00 .. 0.10
    tax = f64_op(f64_of_int(74, 9), b"/", 100.0)                                                      # 0.00 .. 0.08
    return {"l_shipdate": (shipdate, g.I64), "l_quantity": (qty, g.F64), "l_extendedprice": (price, g.F64), "l_discount": (discount, g.F64),
            "l_tax": (tax, g.F64), "l_returnflag": (ctx.generate(n, 0, 75, 3), g.I64), "l_linestatus": (ctx.generate(n, 0, 76, 2), g.I64)}


def tpch_plans(g):
    """(Q6, Q1) of BASELINE config 4 as PhysicalPlan trees over a registered "lineitem" table."""
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
    return q6, q1


def bench_tpch(ctx, g, args, peak):
    """C4: lineitem-shaped columns registered once as a device-resident table; Q6 and Q1 as PhysicalPlan trees through
    olap_execute.  Time = wall clock around execute() (plan lowering, kernels, export of the tiny result)."""
    import time
    n = int(6_000_379 * args.sf)
    ex = g.GpuExecutor(g.EngineConfig(), ctx=ctx)
    ex.drop_table("lineitem")
    cols = device_lineitem(ctx, g, n)
    ex.register_device_table("lineitem", cols, n)
    del cols
    q6, q1 = tpch_plans(g)
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
