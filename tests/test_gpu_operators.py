"""Operator-level parity: PhysicalPlan trees built the way a reference user builds them, executed through the C-ABI
operator layer (olap_execute) on the GPU and compared with the oracle's restatement of executor.rs."""
import numpy as np
import pyarrow as pa
import pytest

import oracle as O
from oracle import executor as OX

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ex(ctx):
    import gpu_olap_engine_b200 as g
    return g.GpuExecutor(g.EngineConfig(), ctx=ctx)


def col(name, index=0):
    import gpu_olap_engine_b200 as g
    return g.Column(index, name)     # the planner always sets index 0 and resolves by name (physical_plan.rs:199)


def lit(v):
    import gpu_olap_engine_b200 as g
    return g.Literal(str(v))


def run_both(ex, plan, tables, sort_by=None, exact=True):
    for name, rbs in tables.items():
        ex.drop_table(name)
        if rbs:
            ex.register_table(name, rbs)
    got = ex.execute(plan)
    exp = OX.execute(plan, tables)
    assert len(got) == len(exp)
    for g_rb, eb in zip(got, exp):
        names, cols, n = eb[:3]
        evalid = [np.ones(n, np.uint8) if v is None else v for v in (eb[3] if len(eb) > 3 else [None] * len(cols))]
        assert g_rb.num_rows == n and g_rb.schema.names == names
        # a NULL is exported as a cleared validity bit over the value 0 (record_batch_convert.rs:103, :108, :161-170)
        gvalid = [np.asarray(g_rb.column(i).is_valid()).astype(np.uint8) for i in range(g_rb.num_columns)]
        gcols = [g_rb.column(i).fill_null(0).to_numpy(zero_copy_only=False) for i in range(g_rb.num_columns)]
        if sort_by is not None and n:
            keys = [names.index(s) for s in sort_by]
            go = np.lexsort([x for k in reversed(keys) for x in (gcols[k], gvalid[k])])
            eo = np.lexsort([x for k in reversed(keys) for x in (cols[k], evalid[k])])
            gcols = [c[go] for c in gcols]; cols = [c[eo] for c in cols]
            gvalid = [v[go] for v in gvalid]; evalid = [v[eo] for v in evalid]
        for name, gc, ec, gv, ev in zip(names, gcols, cols, gvalid, evalid):
            assert gc.dtype == ec.dtype, name
            assert (gv == ev).all(), f"{name}: validity"
            assert (gc == ec).all() if exact else np.allclose(gc, ec, rtol=1e-12, atol=0), name
    return got


def make_sales(n, seed, batches=3):
    k = O.gen_uniform_i64(n, seed, 1000); vi = O.gen_uniform_i64(n, seed + 1, 1 << 20); vf = O.gen_dyadic_f64(n, seed + 2, 1 << 20)
    q = O.gen_uniform_i64(n, seed + 3, 50).astype(np.int32)
    tbl = pa.table({"k": k, "v_i": vi, "v_f": vf, "qty": q})
    edges = np.linspace(0, n, batches + 1).astype(int)
    return [tbl.slice(a, b - a).combine_chunks().to_batches()[0] for a, b in zip(edges[:-1], edges[1:]) if b > a]


def test_reference_filter_unit_vector(ex):
    # gpu-olap-core/src/executor.rs:553-570
    import gpu_olap_engine_b200 as g
    tables = {"t": [pa.record_batch({"a": pa.array([5, 15, 25], pa.int64())})]}
    plan = g.GpuFilter(g.GpuTableScan("t"), g.BinaryExpr(col("a"), ">", lit(10)))
    got = run_both(ex, plan, tables)
    assert got[0].column(0).to_pylist() == [15, 25]


@pytest.mark.parametrize("op", ["=", "==", "!=", "<>", ">", ">=", "<", "<=", "~"])
def test_filter_ops_scan_and_drop_empty(ex, op):
    import gpu_olap_engine_b200 as g
    tables = {"sales": make_sales(30000, 1)}
    run_both(ex, g.GpuFilter(g.GpuTableScan("sales"), g.BinaryExpr(col("qty"), op, lit(7))), tables)
    # literal on the left, column vs column, non-Int64 comparison falls through to "pass all" (executor.rs:471-472)
    run_both(ex, g.GpuFilter(g.GpuTableScan("sales"), g.BinaryExpr(lit(500), op, col("k"))), tables)
    run_both(ex, g.GpuFilter(g.GpuTableScan("sales"), g.BinaryExpr(col("k"), op, col("qty"))), tables)
    run_both(ex, g.GpuFilter(g.GpuTableScan("sales"), g.BinaryExpr(col("k"), op, col("v_f"))), tables)


def test_filter_conjunction_and_f64_extension(ex):
    import gpu_olap_engine_b200 as g
    tables = {"sales": make_sales(50000, 2)}
    pred = g.BinaryExpr(g.BinaryExpr(col("v_i"), ">", lit(1 << 19)), "AND", g.BinaryExpr(g.BinaryExpr(col("k"), "<", lit(700)), "AND",
                                                                                         g.BinaryExpr(col("v_f"), "<=", lit("512.25"))))
    run_both(ex, g.GpuFilter(g.GpuTableScan("sales"), pred), tables)


def test_projection_names_literals_and_arithmetic(ex):
    import gpu_olap_engine_b200 as g
    tables = {"sales": make_sales(20000, 3)}
    exprs = [col("k"), lit(42), g.BinaryExpr(col("v_i"), "+", col("k")), g.BinaryExpr(g.BinaryExpr(col("v_i"), "*", lit(3)), "-", col("qty")),
             g.BinaryExpr(col("v_i"), "/", g.BinaryExpr(col("k"), "-", lit(500))), g.BinaryExpr(col("k"), "%", lit(3)),
             g.BinaryExpr(col("v_f"), "*", lit("0.5"))]
    got = run_both(ex, g.GpuProjection(g.GpuTableScan("sales", projection=[0, 1, 2, 3]), exprs), tables)
    assert got[0].schema.names[:3] == ["k", "lit(42)", "(v_i + k)"]


def test_empty_scan_uses_plan_schema(ex):
    # execute_table_scan (executor.rs:110-155): no data -> one empty batch; float fields Float64, everything else Int64
    import gpu_olap_engine_b200 as g
    schema = pa.schema([("a", pa.int32()), ("b", pa.float32()), ("c", pa.string())])
    got = run_both(ex, g.GpuTableScan("nothing_here", schema=schema, projection=[1, 0]), {"nothing_here": []})
    assert got[0].num_rows == 0 and got[0].schema.types == [pa.float64(), pa.int64()]


def test_global_aggregate_is_the_reference_execution(ex):
    import gpu_olap_engine_b200 as g
    tables = {"sales": make_sales(100000, 4)}
    aggs = [g.Sum(col("v_i")), g.Count(col("v_i")), g.Min(col("v_i")), g.Max(col("v_i")), g.Avg(col("v_i")), g.Sum(col("v_f")), g.Avg(col("qty")),
            g.Sum(g.BinaryExpr(col("v_i"), "*", col("qty")))]
    got = run_both(ex, g.GpuAggregate(g.GpuTableScan("sales"), [], aggs), tables)
    assert all(t == pa.float64() for t in got[0].schema.types) and got[0].num_rows == 1   # executor.rs:347: everything Float64
    # filter that removes every row, then aggregate: SUM 0, COUNT 0, AVG 0, MIN f64::MAX, MAX f64::MIN (executor.rs:310-345)
    plan = g.GpuAggregate(g.GpuFilter(g.GpuTableScan("sales"), g.BinaryExpr(col("k"), "<", lit(-1))), [], aggs[:5])
    got = run_both(ex, plan, tables)
    assert got[0].to_pydict() == {"SUM(v_i)": [0.0], "COUNT(v_i)": [0.0], "MIN(v_i)": [np.finfo(np.float64).max],
                                  "MAX(v_i)": [-np.finfo(np.float64).max], "AVG(v_i)": [0.0]}


def test_filter_group_by_c1a_shape(ex):
    # examples/rust_usage.rs shape (BASELINE config 1): WHERE v_i > 2^19 GROUP BY k  SUM / COUNT / MIN / MAX / AVG
    import gpu_olap_engine_b200 as g
    tables = {"sales": make_sales(1 << 20, 5, batches=4)}
    plan = g.GpuAggregate(g.GpuFilter(g.GpuTableScan("sales"), g.BinaryExpr(col("v_i"), ">", lit(1 << 19))), [col("k")],
                          [g.Sum(col("v_i")), g.Count(col("v_i")), g.Min(col("v_i")), g.Max(col("v_f")), g.Avg(col("v_f")), g.Sum(col("v_f"))])
    got = run_both(ex, plan, tables, sort_by=["k"])
    assert got[0].num_rows == 1000


def test_hash_join_c1b_shape_materialised(ex):
    # BASELINE config 1: 100 K build rows (left) joined with 1 M probe rows; output = left columns then right columns
    import gpu_olap_engine_b200 as g
    nb, npb = 100000, 1000000
    left = pa.table({"id": O.gen_perm_i64(nb, 2, nb), "name_hash": O.gen_uniform_i64(nb, 8, 1 << 40), "w": O.gen_dyadic_f64(nb, 9, 1 << 10)})
    right = pa.table({"cust": O.gen_uniform_i64(npb, 22, nb), "amount": O.gen_uniform_i64(npb, 23, 10000)})
    tables = {"customers": left.to_batches(max_chunksize=40000), "orders": right.to_batches(max_chunksize=300000)}
    plan = g.GpuHashJoin(g.GpuTableScan("customers"), g.GpuTableScan("orders"), [col("id")], [col("cust")], g.JoinType.Inner)
    got = run_both(ex, plan, tables, sort_by=["cust", "amount", "id"])
    assert got[0].num_rows == npb and got[0].schema.names == ["id", "name_hash", "w", "cust", "amount"]
    # join feeding an aggregate (join -> group by), duplicates on the build side, one side empty
    dup = pa.table({"id": O.gen_uniform_i64(5000, 3, 800), "x": O.gen_uniform_i64(5000, 4, 100)})
    tables2 = {"customers": dup.to_batches(), "orders": right.slice(0, 50000).to_batches()}
    plan2 = g.GpuAggregate(g.GpuHashJoin(g.GpuTableScan("customers"), g.GpuTableScan("orders"), [col("id")], [col("cust")]), [col("id")],
                           [g.Count(col("amount")), g.Sum(g.BinaryExpr(col("amount"), "*", col("x")))])
    run_both(ex, plan2, tables2, sort_by=["id"])
    empty = g.GpuFilter(g.GpuTableScan("orders"), g.BinaryExpr(col("cust"), "<", lit(0)))
    assert ex.execute(g.GpuHashJoin(g.GpuTableScan("customers"), empty, [col("id")], [col("cust")])) == []   # executor.rs:231-233


@pytest.mark.parametrize("jt", ["Left", "Right", "Full"])
def test_outer_joins_append_null_extended_rows(ex, jt):
    # JoinType::{Left, Right, Full} (logical_plan.rs:52-58): unmatched rows of the preserved side(s), the other side NULL
    import gpu_olap_engine_b200 as g
    nb, npb = 5000, 40000
    left = pa.table({"id": O.gen_perm_i64(nb, 2, 3 * nb), "w": O.gen_dyadic_f64(nb, 9, 1 << 10)})            # a third of the key domain
    right = pa.table({"cust": O.gen_uniform_i64(npb, 22, 4 * nb), "amount": O.gen_uniform_i64(npb, 23, 10000)})  # some keys match nothing
    tables = {"customers": left.to_batches(max_chunksize=2000), "orders": right.to_batches(max_chunksize=15000)}
    plan = g.GpuHashJoin(g.GpuTableScan("customers"), g.GpuTableScan("orders"), [col("id")], [col("cust")], getattr(g.JoinType, jt))
    got = run_both(ex, plan, tables, sort_by=["cust", "amount", "id", "w"])
    rb = got[0]
    lk, rk = left["id"].to_numpy(), right["cust"].to_numpy()
    inner = int(np.isin(rk, lk).sum())                      # unique build keys: one pair per matching probe row
    ul = int((~np.isin(lk, rk)).sum()) if jt in ("Left", "Full") else 0
    ur = int((~np.isin(rk, lk)).sum()) if jt in ("Right", "Full") else 0
    assert ul + ur > 0 and rb.num_rows == inner + ul + ur
    assert rb.column("cust").null_count == ul and rb.column("amount").null_count == ul
    assert rb.column("id").null_count == ur and rb.column("w").null_count == ur
    # the unmatched rows come behind the inner pairs, in row order
    if ul:
        tail = rb.column("id").to_numpy(zero_copy_only=False)[inner:inner + ul].astype(np.int64)
        assert (tail == lk[~np.isin(lk, rk)]).all()
    if ur:
        tail = rb.column("cust").to_numpy(zero_copy_only=False)[inner + ul:].astype(np.int64)
        assert (tail == rk[~np.isin(rk, lk)]).all()


def test_outer_joins_edge_cases(ex):
    import gpu_olap_engine_b200 as g
    # duplicates on both sides, nothing matches at all, everything matches
    l = pa.table({"a": np.array([1, 1, 2, 3, 7], np.int64), "x": np.arange(5, dtype=np.int64)})
    r = pa.table({"b": np.array([1, 2, 2, 9], np.int64), "y": np.arange(4, dtype=np.float64)})
    for jt in (g.JoinType.Inner, g.JoinType.Left, g.JoinType.Right, g.JoinType.Full):
        plan = g.GpuHashJoin(g.GpuTableScan("l"), g.GpuTableScan("r"), [col("a")], [col("b")], jt)
        got = run_both(ex, plan, {"l": l.to_batches(), "r": r.to_batches()}, sort_by=["a", "x", "b", "y"])
        assert got[0].num_rows == {0: 4, 1: 6, 2: 5, 3: 7}[int(jt)]
    disjoint = pa.table({"b": np.array([100, 200], np.int64), "y": np.array([0.5, 1.5])})
    full = run_both(ex, g.GpuHashJoin(g.GpuTableScan("l"), g.GpuTableScan("r"), [col("a")], [col("b")], g.JoinType.Full),
                    {"l": l.to_batches(), "r": disjoint.to_batches()}, sort_by=["a", "x", "b", "y"])
    assert full[0].num_rows == 7 and full[0].column("a").null_count == 2 and full[0].column("y").null_count == 5
    same = run_both(ex, g.GpuHashJoin(g.GpuTableScan("l"), g.GpuTableScan("l"), [col("a")], [col("a")], g.JoinType.Full),
                    {"l": l.to_batches()}, sort_by=["a", "x"])
    assert same[0].num_rows == 7 and all(same[0].column(i).null_count == 0 for i in range(4))
    # the kernel seam under it: rows of a side that appear in no pair, ascending
    pairs = ex.ctx.upload_raw(np.array([[0, 3], [2, 3], [2, 0]], np.uint32).reshape(-1))
    assert ex.ctx.join_unmatched_rows(pairs, 3, 0, 5).tolist() == [1, 3, 4] and ex.ctx.join_unmatched_rows(pairs, 3, 1, 4).tolist() == [1, 2]
    assert ex.ctx.join_unmatched_rows(None, 0, 1, 3).tolist() == [0, 1, 2] and ex.ctx.join_unmatched_rows(pairs, 3, 0, 0).tolist() == []
    # ... and the materialisation of a payload column in pair order (olap_join_gather_8b), both sides, odd sizes
    rng = np.random.default_rng(3)
    for n_pairs in (3, 1000, 262_147):
        pr = rng.integers(0, 5000, (n_pairs, 2)).astype(np.uint32)
        colv = rng.integers(-2**62, 2**62, 5000).astype(np.int64)
        dpairs, dcol, dout = ex.ctx.upload_raw(pr.reshape(-1)), ex.ctx.upload_raw(colv), ex.ctx.alloc(n_pairs, np.int64)
        for side in (0, 1):
            g._lib.check(ex.ctx.lib.olap_join_gather_8b(ex.ctx.handle, dcol.ptr, dpairs.ptr, side, n_pairs, dout.ptr, None))
            assert (dout.to_numpy() == colv[pr[:, side]]).all()


def test_arrow_import_layout_rules(ex):
    # record_batch_convert.rs:35-130: widening, u64 reinterpretation, nulls -> 0, Utf8 -> FNV-1a
    import gpu_olap_engine_b200 as g
    rb = pa.record_batch({"i8": pa.array([1, None, -3], pa.int8()), "u64": pa.array([2**64 - 1, 5, None], pa.uint64()),
                          "f32": pa.array([1.5, None, -2.25], pa.float32()), "s": pa.array(["a", None, "foobar"])})
    got = run_both(ex, g.GpuTableScan("mixed"), {"mixed": [rb]})
    # nulls: the value slot holds 0 (:103, :108), the validity byte brings the NULL back on the way out (buffer_to_array, :161-170)
    assert got[0].column("i8").to_pylist() == [1, None, -3] and got[0].column("u64").to_pylist() == [-1, 5, None]
    assert got[0].column("i8").fill_null(0).to_pylist() == [1, 0, -3] and got[0].column("f32").to_pylist() == [1.5, None, -2.25]
    # ... and every operator downstream reads the 0: filters / expressions emit non-null columns (take_int64, executor.rs:496-507)
    kept = run_both(ex, g.GpuFilter(g.GpuTableScan("mixed"), g.BinaryExpr(col("i8"), ">=", lit(0))), {"mixed": [rb]})
    assert kept[0].column("i8").to_pylist() == [1, 0] and kept[0].column("i8").null_count == 0
    proj = run_both(ex, g.GpuProjection(g.GpuTableScan("mixed"), [col("u64"), g.BinaryExpr(col("i8"), "+", lit(1))]), {"mixed": [rb]})
    assert proj[0].column(0).to_pylist() == [-1, 5, None] and proj[0].column(1).to_pylist() == [2, 1, -2]
    assert got[0].column("s").to_pylist()[2] == np.uint64(0x85944171f73967e8).astype(np.int64)
    sl = pa.record_batch({"a": pa.array(range(100), pa.int16())}).slice(10, 50)     # non-zero Arrow offset
    run_both(ex, g.GpuTableScan("sliced"), {"sliced": [sl]})


def test_operator_errors(ex):
    import gpu_olap_engine_b200 as g
    ex.drop_table("sales"); ex.register_table("sales", make_sales(1000, 6))
    with pytest.raises(g.OlapError, match="not found"):
        ex.execute(g.GpuProjection(g.GpuTableScan("sales"), [g.Column(99, "nope")]))
    with pytest.raises(g.OlapError, match="Cannot parse literal"):
        ex.execute(g.GpuProjection(g.GpuTableScan("sales"), [lit("abc")]))
    with pytest.raises(g.OlapError, match="not registered"):
        ex.execute(g.GpuTableScan("missing"))


def make_lineitem(n, seed):
    """Synthetic lineitem columns of SURVEY.md 8(d) C4 (values are cents / hundredths so f64 products stay exact enough)."""
    u = lambda s, d: O.gen_uniform_i64(n, seed + s, d)
    qty = (u(1, 50) + 1).astype(np.float64)
    price = np.round(qty * (900.0 + u(2, 120001) / 100.0) * 100.0) / 100.0
    return pa.table({"l_shipdate": u(0, 2526) + 8036, "l_quantity": qty, "l_extendedprice": price, "l_discount": u(3, 11) / 100.0,
                     "l_tax": u(4, 9) / 100.0, "l_returnflag": u(5, 3), "l_linestatus": u(6, 2)})


def test_tpch_q6_shape(ex):
    # BASELINE config 4 (Q6): shipdate in [1994-01-01, 1995-01-01) AND discount in [0.05, 0.07] AND quantity < 24;
    # SUM(extendedprice * discount).  Float64 predicates / arithmetic are this core's extension.
    import gpu_olap_engine_b200 as g
    t = make_lineitem(400000, 50)
    tables = {"lineitem": t.to_batches(max_chunksize=150000)}
    AND = lambda a, b: g.BinaryExpr(a, "AND", b)
    pred = AND(AND(g.BinaryExpr(col("l_shipdate"), ">=", lit(8766)), g.BinaryExpr(col("l_shipdate"), "<", lit(9131))),
               AND(AND(g.BinaryExpr(col("l_discount"), ">=", lit("0.05")), g.BinaryExpr(col("l_discount"), "<=", lit("0.07"))),
                   g.BinaryExpr(col("l_quantity"), "<", lit("24.0"))))
    plan = g.GpuAggregate(g.GpuFilter(g.GpuTableScan("lineitem"), pred), [], [g.Sum(g.BinaryExpr(col("l_extendedprice"), "*", col("l_discount"))),
                                                                           g.Count(col("l_quantity"))])
    got = run_both(ex, plan, tables, exact=False)      # f64 SUM over random data: <= 1e-12 relative (order differs)
    assert got[0].column(1)[0].as_py() > 1000


def test_tpch_q1_shape(ex):
    # BASELINE config 4 (Q1): shipdate <= 1998-09-02, GROUP BY returnflag * 2 + linestatus, 8 aggregates
    import gpu_olap_engine_b200 as g
    t = make_lineitem(300000, 60)
    tables = {"lineitem": t.to_batches(max_chunksize=100000)}
    one = lit("1.0")
    disc_price = g.BinaryExpr(col("l_extendedprice"), "*", g.BinaryExpr(one, "-", col("l_discount")))
    charge = g.BinaryExpr(disc_price, "*", g.BinaryExpr(one, "+", col("l_tax")))
    key = g.BinaryExpr(g.BinaryExpr(col("l_returnflag"), "*", lit(2)), "+", col("l_linestatus"))
    aggs = [g.Sum(col("l_quantity")), g.Sum(col("l_extendedprice")), g.Sum(disc_price), g.Sum(charge), g.Avg(col("l_quantity")),
            g.Avg(col("l_extendedprice")), g.Avg(col("l_discount")), g.Count(col("l_quantity"))]
    plan = g.GpuAggregate(g.GpuFilter(g.GpuTableScan("lineitem"), g.BinaryExpr(col("l_shipdate"), "<=", lit(10471))), [key], aggs)
    got = run_both(ex, plan, tables, sort_by=["((l_returnflag * lit(2)) + l_linestatus)"], exact=False)
    assert got[0].num_rows == 6


def test_aggregate_expression_inputs_fused_and_materialised(ex):
    # SUM / MIN / AVG over arithmetic expressions: with few groups the expression is evaluated inside the aggregation
    # kernel (olap_group_by_prog_i64); with 300 K groups over 2^21 rows the kernel refuses (partitioned regime) and the
    # operator layer materialises the expression instead -- same answers either way
    import gpu_olap_engine_b200 as g
    n = (1 << 21) + 123
    a = O.gen_uniform_i64(n, 71, 1 << 20) - (1 << 19); b = O.gen_uniform_i64(n, 72, 1000) + 1
    f = O.gen_dyadic_f64(n, 73, 1 << 16)
    for groups in (7, 300000):
        t = pa.table({"k": O.gen_uniform_i64(n, 70, groups), "a": a, "b": b, "f": f})
        tables = {"t": t.to_batches(max_chunksize=700000)}
        e1 = g.BinaryExpr(g.BinaryExpr(col("a"), "*", lit(3)), "-", g.BinaryExpr(col("a"), "/", col("b")))
        e2 = g.BinaryExpr(col("f"), "*", g.BinaryExpr(lit("0.5"), "+", col("f")))
        plan = g.GpuAggregate(g.GpuTableScan("t"), [col("k")], [g.Sum(e1), g.Min(e1), g.Avg(e1), g.Sum(e2), g.Max(e2), g.Count(col("a"))])
        run_both(ex, plan, tables, sort_by=["k"], exact=True)

