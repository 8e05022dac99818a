"""CPU restatement of GpuExecutor::execute (gpu-olap-core/src/executor.rs:37-355) over numpy columns, built from the
oracle's primitives.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Stubs are completed per SURVEY.md 8(c).

A batch is (names, [np.ndarray int64|float64], n_rows, valids); valids[i] is None (no nulls) or a uint8 array, 1 = valid
(column_buffer.rs:37-41; buffer_to_array puts the nulls back, record_batch_convert.rs:161-170).  Scans carry validity, outer joins
produce it; filters, expressions, join keys and aggregates read a null as the 0 stored in its place (:103, :108) and emit
non-null columns, as take_int64 / eval_expr do (executor.rs:496-507, :377-409).  tables is {name: [pyarrow.RecordBatch]}."""
from __future__ import annotations

import numpy as np

import oracle as O

_CMP = {"=": O.EQ, "==": O.EQ, "!=": O.NE, "<>": O.NE, ">": O.GT, ">=": O.GE, "<": O.LT, "<=": O.LE}


def _widen_batch(rb):
    import pyarrow as pa
    names, cols, valids = [], [], []
    for f, c in zip(rb.schema, rb.columns):
        names.append(f.name)
        valid = None if c.null_count == 0 else np.array([v is not None for v in c.to_pylist()], np.uint8)
        valids.append(valid)
        if pa.types.is_string(f.type) or pa.types.is_large_string(f.type):   # record_batch_convert.rs:93-97, :112-121
            cols.append(np.array([0 if v is None else np.uint64(O.fnv1a64(v.encode())).astype(np.int64) for v in c.to_pylist()], np.int64))
        else:
            raw = np.asarray(c.fill_null(0)) if c.null_count else c.to_numpy(zero_copy_only=False)
            cols.append(O.widen_column(np.ascontiguousarray(raw), validity=valid))
    return names, cols, rb.num_rows, valids


def _valids(b):
    return b[3] if len(b) > 3 and b[3] is not None else [None] * len(b[1])


def expr_name(e):                                     # executor.rs:529-537
    k = type(e).__name__
    if k == "Column":
        return e.name
    if k == "Literal":
        return f"lit({e.value})"
    return f"({expr_name(e.left)} {e.op} {expr_name(e.right)})"


def eval_expr(e, batch):                              # executor.rs:377-441
    names, cols, n = batch[:3]
    k = type(e).__name__
    if k == "Column":
        if e.name in names:
            return cols[names.index(e.name)]
        if e.index < len(cols):
            return cols[e.index]
        raise KeyError(f"Column '{e.name}' not found in batch")
    if k == "Literal":
        try:
            return np.full(n, int(e.value), np.int64)
        except ValueError:
            return np.full(n, float(e.value), np.float64)
    l, r = eval_expr(e.left, batch), eval_expr(e.right, batch)
    if l.dtype != r.dtype:
        raise TypeError("Binary expr: mixed operand types")
    if l.dtype == np.int64:
        return O.project_i64(l, e.op if len(e.op) == 1 else "?", r)
    with np.errstate(all="ignore"):                   # Float64 arithmetic: this core's extension
        return {"+": l + r, "-": l - r, "*": l * r, "/": l / r}.get(e.op, np.zeros(n))


def eval_filter_mask(batch, pred):                    # executor.rs:443-476 (+ real AND, + f64 compare extension)
    n = batch[2]
    if type(pred).__name__ != "BinaryExpr":
        return np.ones(n, np.uint8)
    if pred.op in ("AND", "and", "&&"):
        return eval_filter_mask(batch, pred.left) & eval_filter_mask(batch, pred.right)
    l, r = eval_expr(pred.left, batch), eval_expr(pred.right, batch)
    if l.dtype != r.dtype:
        return np.ones(n, np.uint8)
    op = _CMP.get(pred.op, 99)
    if l.dtype == np.int64:
        return O.filter_mask_i64(l, op, rhs=r)
    f = {O.EQ: l == r, O.NE: l != r, O.LT: l < r, O.LE: l <= r, O.GT: l > r, O.GE: l >= r}.get(op, np.zeros(n, bool))
    return f.astype(np.uint8)


def execute(plan, tables):
    k = type(plan).__name__
    if k == "GpuTableScan":
        rbs = tables.get(plan.table_name, [])
        if not rbs:
            import pyarrow as pa
            fields = list(plan.schema)
            names = [f.name for f in fields]
            cols = [np.empty(0, np.float64 if pa.types.is_floating(f.type) else np.int64) for f in fields]
            out = [(names, cols, 0, [None] * len(cols))]
        else:
            out = [_widen_batch(rb) for rb in rbs]
        if plan.projection is not None:
            out = [([b[0][i] for i in plan.projection], [b[1][i] for i in plan.projection], b[2], [_valids(b)[i] for i in plan.projection]) for b in out]
        return out
    if k == "GpuFilter":
        out = []
        for b in execute(plan.input, tables):
            idx = O.mask_to_indices(eval_filter_mask(b, plan.predicate))
            if len(idx):                              # executor.rs:173
                out.append((b[0], [O.take(c, idx) for c in b[1]], len(idx)))
        return out
    if k == "GpuProjection":
        def passthrough(e, b):                        # a bare column keeps its validity
            if type(e).__name__ != "Column":
                return None
            v = _valids(b)
            return v[b[0].index(e.name)] if e.name in b[0] else (v[e.index] if e.index < len(v) else None)
        return [([expr_name(e) for e in plan.exprs], [eval_expr(e, b) for e in plan.exprs], b[2], [passthrough(e, b) for e in plan.exprs])
                for b in execute(plan.input, tables)]
    if k == "GpuHashJoin":
        left, right = execute(plan.left, tables), execute(plan.right, tables)
        if not left or not right:                     # executor.rs:231-233
            return []
        lk = np.concatenate([eval_expr(plan.left_keys[0], b) for b in left])
        rk = np.concatenate([eval_expr(plan.right_keys[0], b) for b in right])
        cnt, pairs = O.hash_join_i64(lk, rk)
        names = left[0][0] + right[0][0]
        lcols = [np.concatenate([b[1][i] for b in left]) for i in range(len(left[0][0]))]
        rcols = [np.concatenate([b[1][i] for b in right]) for i in range(len(right[0][0]))]
        jt = int(getattr(plan, "join_type", 0))      # JoinType (logical_plan.rs:52-58): 0 Inner, 1 Left, 2 Right, 3 Full
        # outer joins (never executed by the reference; SQL semantics): the preserved side's rows without a partner follow the inner
        # pairs in row order -- unmatched left rows, then unmatched right rows -- with the other side NULL (value 0, validity 0)
        ul = np.setdiff1d(np.arange(len(lk)), pairs[:, 0]) if jt in (1, 3) else np.empty(0, np.int64)
        ur = np.setdiff1d(np.arange(len(rk)), pairs[:, 1]) if jt in (2, 3) else np.empty(0, np.int64)
        cols, valids = [], []
        for c in lcols:
            cols.append(np.concatenate([c[pairs[:, 0]], c[ul], np.zeros(len(ur), c.dtype)]))
            valids.append(np.concatenate([np.ones(cnt + len(ul), np.uint8), np.zeros(len(ur), np.uint8)]) if len(ur) else None)
        for c in rcols:
            cols.append(np.concatenate([c[pairs[:, 1]], np.zeros(len(ul), c.dtype), c[ur]]))
            valids.append(np.concatenate([np.ones(cnt, np.uint8), np.zeros(len(ul), np.uint8), np.ones(len(ur), np.uint8)]) if len(ul) else None)
        return [(names, cols, cnt + len(ul) + len(ur), valids)]
    if k == "GpuAggregate":
        inp = execute(plan.input, tables)
        names_of = {0: "SUM", 1: "COUNT", 2: "MIN", 3: "MAX", 4: "AVG"}
        names = [f"{names_of[a.func]}({expr_name(a.input)})" for a in plan.aggr_exprs]
        if not plan.group_by:                         # the reference's actual execution: executor.rs:308-349
            cols = []
            for a in plan.aggr_exprs:
                st = O.GlobalAgg()
                for b in inp:
                    st.update(eval_expr(a.input, b))
                cols.append(np.array([st.final(a.func)], np.float64))
            return [(names, cols, 1)]
        if not inp or sum(b[2] for b in inp) == 0:
            return []
        keys = np.concatenate([eval_expr(plan.group_by[0], b) for b in inp])
        vals = [np.concatenate([eval_expr(a.input, b) for b in inp]) for a in plan.aggr_exprs]
        gk, res = O.group_by_i64(keys, vals, [a.func for a in plan.aggr_exprs])
        return [([expr_name(plan.group_by[0])] + names, [gk] + res, len(gk))]
    raise TypeError(k)
