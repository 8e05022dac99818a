"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle finishes these sizes in minutes,
not seconds, so the checks are: exact counts, permutation / key-equality of the join pairs, checksums of checksums across
independent regimes, sortedness of selection vectors).  All calls go through the C ABI; torch is used only to look at the
device results (sums, scatter, comparisons) without a 4 GB download.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu


class _Cai:
    """Expose a DeviceBuffer region to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr, "version": 3}


def as_torch(buf, shape, typestr):
    return torch.as_tensor(_Cai(buf.ptr, shape, typestr), device="cuda")


def test_join_c2_full_size_properties(ctx):
    # BASELINE configs[1]: 2^27 unique build keys x 2^29 FK probes, every probe row matches exactly one build row
    nb, npb = 1 << 27, 1 << 29
    db = ctx.generate(nb, 1, 3, nb)          # a permutation of 0 .. nb-1
    dp = ctx.generate(npb, 0, 33, nb)        # uniform over the build keys
    cnt, pairs = ctx.hash_join_i64(db, nb, dp, npb, max_matches=npb, strategy=0)
    ctx.synchronize()
    assert cnt == npb
    pr = as_torch(pairs, (npb, 2), "<i4")    # row ids < 2^31
    bk = as_torch(db, (nb,), "<i8"); pk = as_torch(dp, (npb,), "<i8")
    left, right = pr[:, 0].long(), pr[:, 1].long()
    assert int(left.min()) >= 0 and int(left.max()) < nb and int(right.min()) >= 0 and int(right.max()) < npb
    seen = torch.zeros(npb, dtype=torch.bool, device="cuda")
    seen[right] = True
    assert bool(seen.all())                                   # every probe row appears (count == npb: exactly once)
    assert bool((bk[left] == pk[right]).all())               # with an equal build key
    del seen, left, right, pr, bk, pk
    torch.cuda.empty_cache()


def test_group_by_full_size_checksums(ctx):
    # 2^30 rows: the partitioned regime (1e5 groups) against the register (no key) regime and torch's own reductions
    import gpu_olap_engine_b200 as g
    n, G = 1 << 30, 100000
    dk = ctx.generate(n, 0, 4, G)
    dv = ctx.generate(n, 0, 5, 1 << 20)
    aggs = [(g.SUM, dv, g.I64), (g.COUNT, dv, g.I64), (g.MIN, dv, g.I64), (g.MAX, dv, g.I64)]
    ng, ok, oa, cap = ctx.group_by_i64(dk, n, aggs, G)
    ctx.synchronize()
    assert ng == G
    keys = as_torch(ok, (ng,), "<i8"); out = as_torch(oa, (4, cap), "<i8")[:, :ng]
    v = as_torch(dv, (n,), "<i8")
    assert int(keys.min()) == 0 and int(keys.max()) == G - 1 and int(torch.unique(keys).numel()) == G
    assert int(out[1].sum()) == n                              # counts add up to the rows
    assert int(out[0].sum()) == int(v.sum())                   # sums add up to the column sum
    assert int(out[2].min()) == int(v.min()) and int(out[3].max()) == int(v.max())
    # spot check of four groups against torch, row by row
    k = as_torch(dk, (n,), "<i8")
    order = torch.argsort(keys)
    for gi in (0, 1, 777, G - 1):
        m = k == gi
        row = int(order[gi])
        assert int(out[1, row]) == int(m.sum()) and int(out[0, row]) == int(v[m].sum())
        assert int(out[2, row]) == int(v[m].min()) and int(out[3, row]) == int(v[m].max())
    ng1, ok1, oa1, cap1 = ctx.group_by_i64(None, n, aggs, 1)
    ctx.synchronize()
    tot = as_torch(oa1, (4, cap1), "<i8")[:, 0]
    assert ng1 == 1 and int(tot[0]) == int(out[0].sum()) and int(tot[1]) == n
    del keys, out, v, k, tot
    torch.cuda.empty_cache()


def test_filter_full_size_selection_vector(ctx):
    import gpu_olap_engine_b200 as g
    n = 1 << 30
    dv = ctx.generate(n, 0, 9, 1000)
    cnt, mask, sel = ctx.filter([g.Term(dv, g.GT, 499)], n)
    ctx.synchronize()
    v = as_torch(dv, (n,), "<i8")
    exp = v > 499
    assert cnt == int(exp.sum())
    assert bool((as_torch(mask, (n,), "|u1") == exp.to(torch.uint8)).all())
    s = as_torch(sel, (cnt,), "<i4").long()                    # row ids < 2^31
    assert bool((s[1:] > s[:-1]).all())                        # row order preserved
    assert bool(exp[s].all())                                  # cnt distinct passing rows == all passing rows
    del v, exp, s
    torch.cuda.empty_cache()


def test_tpch_shapes_2_26_rows_device_resident(ctx):
    # BASELINE config 4 shapes at 2^26 rows through the operator API (olap_execute) on a device-registered table, against torch's own
    # reductions over the same device columns: Q6 (5 fused conjuncts, SUM(price * discount)) and Q1 (GROUP BY returnflag * 2 + linestatus
    # as a fused key program, 8 aggregates over fused value programs).  Integer results exact, f64 sums within 1e-12 relative.
    import gpu_olap_engine_b200 as g
    from bench_ops import device_lineitem, tpch_plans
    n = 1 << 26
    ex = g.GpuExecutor(g.EngineConfig(), ctx=ctx)
    ex.drop_table("lineitem")
    cols = device_lineitem(ctx, g, n)
    ex.register_device_table("lineitem", cols, n)
    t = {k: as_torch(b, (n,), "<i8" if ty == g.I64 else "<f8") for k, (b, ty) in cols.items()}
    q6, q1 = tpch_plans(g)
    ship, qty, price, disc, tax = t["l_shipdate"], t["l_quantity"], t["l_extendedprice"], t["l_discount"], t["l_tax"]
    assert int(ship.min()) == 8036 and int(ship.max()) == 10561 and float(qty.min()) == 1.0 and float(qty.max()) == 50.0
    m6 = (ship >= 8766) & (ship < 9131) & (disc >= 0.05) & (disc <= 0.07) & (qty < 24.0)
    exp6 = float((price * disc)[m6].sum(dtype=torch.float64))
    got6 = ex.execute(q6)
    assert len(got6) == 1 and got6[0].num_rows == 1
    v6 = got6[0].column(0).to_pylist()[0]
    assert abs(v6 - exp6) <= 1e-12 * abs(exp6), (v6, exp6)
    m1 = ship <= 10471
    key = t["l_returnflag"] * 2 + t["l_linestatus"]
    got1 = ex.execute(q1)[0]
    assert got1.num_rows == 6
    gk = np.asarray(got1.column(0).to_numpy())
    order = np.argsort(gk)
    assert gk[order].tolist() == [0, 1, 2, 3, 4, 5]
    dp = price * (1.0 - disc)
    ch = dp * (1.0 + tax)
    # the reference sums are torch.sum over each group's rows (pairwise summation: ~1e-15 relative), not atomics in arbitrary order
    masks = [m1 & (key == k) for k in range(6)]
    cnt = torch.stack([m.sum() for m in masks]).to(torch.float64)
    gsum = lambda x: torch.stack([x[m].sum(dtype=torch.float64) for m in masks])
    exp = [gsum(qty), gsum(price), gsum(dp), gsum(ch)]
    exp += [exp[0] / cnt, exp[1] / cnt, gsum(disc) / cnt, cnt]
    for i, e in enumerate(exp):
        got = np.asarray(got1.column(1 + i).to_numpy(), dtype=np.float64)[order]
        ref = e.cpu().numpy()
        assert np.allclose(got, ref, rtol=1e-12, atol=0), (i, float(np.max(np.abs(got - ref) / np.abs(ref))))
    assert np.asarray(got1.column(8).to_numpy())[order].astype(np.int64).sum() == int(m1.sum())     # COUNT is exact
    ex.drop_table("lineitem")
    del t, cols
    torch.cuda.empty_cache()
