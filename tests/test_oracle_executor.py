"""CPU tests of oracle/executor.py where it completes what the reference leaves out: the outer joins
(JoinType::{Left, Right, Full}, gpu-olap-core/src/logical_plan.rs:52-58 -- never executed by the reference, so "parity unpinned")
are cross-checked against pandas' merge, and NULLs follow record_batch_convert.rs (:103, :108 value 0, :161-170 null on the way out)."""
import numpy as np
import pandas as pd
import pyarrow as pa
import pytest

import oracle as O
from oracle import executor as OX


def _plan(jt):
    from gpu_olap_engine_b200.plan import Column, GpuHashJoin, GpuTableScan, JoinType
    return GpuHashJoin(GpuTableScan("l"), GpuTableScan("r"), [Column(0, "a")], [Column(0, "b")], getattr(JoinType, jt))


def _rows(names, cols, valids):
    out = []
    for i in range(len(cols[0]) if cols else 0):
        out.append(tuple(None if (v is not None and not v[i]) else c[i].item() for c, v in zip(cols, valids)))
    return sorted(out, key=lambda t: tuple((x is None, 0 if x is None else x) for x in t))


@pytest.mark.parametrize("jt,how", [("Inner", "inner"), ("Left", "left"), ("Right", "right"), ("Full", "outer")])
def test_outer_joins_against_pandas(jt, how):
    nb, npb = 300, 2000
    l = pa.table({"a": O.gen_uniform_i64(nb, 2, 400), "x": O.gen_uniform_i64(nb, 3, 1000)})        # duplicates on both sides
    r = pa.table({"b": O.gen_uniform_i64(npb, 4, 600), "y": O.gen_dyadic_f64(npb, 5, 1 << 10)})    # keys 400..599 match nothing
    (names, cols, n, valids), = OX.execute(_plan(jt), {"l": l.to_batches(max_chunksize=100), "r": r.to_batches(max_chunksize=700)})
    assert names == ["a", "x", "b", "y"] and all(len(c) == n for c in cols)
    exp = pd.merge(l.to_pandas(), r.to_pandas(), left_on="a", right_on="b", how=how)
    erows = sorted((tuple(None if pd.isna(v) else (float(v) if k == "y" else int(v)) for k, v in zip(names, row)) for row in exp.itertuples(index=False)),
                   key=lambda t: tuple((x is None, 0 if x is None else x) for x in t))
    assert _rows(names, cols, [None if v is None else v for v in valids]) == erows
    # layout rule: inner pairs first, then unmatched left rows in row order, then unmatched right rows in row order; NULL slots hold 0
    lk, rk = l["a"].to_numpy(), r["b"].to_numpy()
    inner = int(sum((lk == k).sum() for k in rk))
    ul = lk[~np.isin(lk, rk)] if jt in ("Left", "Full") else np.empty(0, np.int64)
    ur = rk[~np.isin(rk, lk)] if jt in ("Right", "Full") else np.empty(0, np.int64)
    assert n == inner + len(ul) + len(ur)
    assert (cols[0][inner:inner + len(ul)] == ul).all() and (cols[2][inner + len(ul):] == ur).all()
    if len(ul):
        assert (cols[2][inner:inner + len(ul)] == 0).all() and (valids[2][inner:inner + len(ul)] == 0).all()
    if len(ur):
        assert (cols[1][inner + len(ul):] == 0).all() and (valids[0][inner + len(ul):] == 0).all()


def test_scan_carries_validity_and_filters_drop_it():
    from gpu_olap_engine_b200.plan import BinaryExpr, Column, GpuFilter, GpuProjection, GpuTableScan, Literal
    rb = pa.record_batch({"i8": pa.array([1, None, -3], pa.int8()), "f": pa.array([None, 2.5, 1.0], pa.float32())})
    (names, cols, n, valids), = OX.execute(GpuTableScan("t"), {"t": [rb]})
    assert cols[0].tolist() == [1, 0, -3] and valids[0].tolist() == [1, 0, 1] and cols[1].tolist() == [0.0, 2.5, 1.0] and valids[1].tolist() == [0, 1, 1]
    out, = OX.execute(GpuFilter(GpuTableScan("t"), BinaryExpr(Column(0, "i8"), ">=", Literal("0"))), {"t": [rb]})
    assert out[1][0].tolist() == [1, 0] and len(out) == 3           # take_int64 (executor.rs:496-507): values only, no nulls
    p, = OX.execute(GpuProjection(GpuTableScan("t"), [Column(0, "i8"), BinaryExpr(Column(0, "i8"), "+", Literal("1"))]), {"t": [rb]})
    assert p[3][0].tolist() == [1, 0, 1] and p[3][1] is None and p[1][1].tolist() == [2, 1, -2]
