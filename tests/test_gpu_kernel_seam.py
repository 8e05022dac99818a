"""The reference's kernel-wrapper seam (gpu-kernel-compiler/src/{filter,join,aggregate}_kernel.rs) driven the way its callers drive it:
raw u64 device addresses in param structs, results through out-pointers -- FilterKernel::launch (filter_kernel.rs:42-72),
JoinKernel::join (join_kernel.rs:39-68), AggregateKernel::aggregate (aggregate_kernel.rs:80-125) -- against the oracle."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


def test_filter_kernel_launch_writes_the_u8_mask(ctx):
    from gpu_olap_engine_b200.kernels import CmpOp, FilterKernel, FilterParams, FilterPredicate
    col = np.array([5, 15, 25], np.int64)                                   # executor.rs:553-570's vector through the kernel seam
    d, mask = ctx.upload_raw(col), ctx.alloc(3, np.uint8)
    grid, block = FilterKernel.compute_launch_params(3)
    kept = FilterKernel(ctx).launch(FilterPredicate(0, CmpOp.Gt, 10), FilterParams(d.ptr, 3, mask.ptr, grid, block))
    assert kept == 2 and mask.to_numpy().tolist() == [0, 1, 1]              # 1 = pass (filter_kernel.rs:3-5)
    n = 300_001
    v = O.gen_uniform_i64(n, 11, 1 << 20)
    d, mask = ctx.upload_raw(v), ctx.alloc(n, np.uint8)
    for op, oop in ((CmpOp.Eq, O.EQ), (CmpOp.Ne, O.NE), (CmpOp.Lt, O.LT), (CmpOp.Le, O.LE), (CmpOp.Gt, O.GT), (CmpOp.Ge, O.GE)):
        kept = FilterKernel(ctx).launch(FilterPredicate(0, op, 1 << 19), FilterParams(d.ptr, n, mask.ptr))
        exp = O.filter_mask_i64(v, oop, literal=1 << 19)
        assert kept == int(exp.sum()) and (mask.to_numpy() == exp).all()


def test_join_kernel_join_fills_pairs_and_count(ctx):
    import gpu_olap_engine_b200 as g
    from gpu_olap_engine_b200.kernels import JoinKernel, JoinParams, JoinStrategy
    nb, npb = 60_000, 250_000
    b, p = O.gen_perm_i64(nb, 2, nb), O.gen_uniform_i64(npb, 22, nb + nb // 4)     # a fifth of the probe keys match nothing
    db, dp = ctx.upload_raw(b), ctx.upload_raw(p)
    ec, epairs = O.hash_join_i64(b, p)
    pairs, cnt = ctx.alloc(2 * npb, np.uint32), ctx.alloc(1, np.uint64)
    jk = JoinKernel(ctx)
    for strategy in (jk.choose_strategy(nb), JoinStrategy.RadixHashJoin):          # Broadcast below 1 M build rows (join_kernel.rs:71-77)
        c = jk.join(strategy, JoinParams(db.ptr, dp.ptr, nb, npb, pairs.ptr, cnt.ptr, max_matches=npb))
        assert c == ec == int(cnt.to_numpy()[0])                                   # out_count_ptr carries the same exact count
        assert (O.sort_pairs(pairs.to_numpy(2 * c).reshape(-1, 2)) == O.sort_pairs(epairs)).all()
    with pytest.raises(g.OlapError) as ei:                                         # the reference truncates silently (join_kernels.cuh:155-160)
        jk.join(JoinStrategy.BroadcastHashJoin, JoinParams(db.ptr, dp.ptr, nb, npb, pairs.ptr, cnt.ptr, max_matches=1000))
    assert ei.value.status == 3 and int(cnt.to_numpy()[0]) == ec                   # ... here it is reported, the count stays exact


def test_aggregate_kernel_aggregate_agg_major_slots(ctx):
    import gpu_olap_engine_b200 as g
    from gpu_olap_engine_b200.kernels import AggFunc, AggregateKernel, AggregateParams, AggregateSpec
    n, G = 200_000, 777
    k, vi, vf = O.gen_uniform_i64(n, 4, G), O.gen_uniform_i64(n, 5, 1 << 20), O.gen_dyadic_f64(n, 6, 1 << 20)
    dk, di, df = ctx.upload_raw(k), ctx.upload_raw(vi), ctx.upload_raw(vf)
    cap = 1024                                                                      # n_groups is the caller's upper bound (aggregate_kernel.rs:52)
    out, out_keys = ctx.alloc(4 * cap, np.uint64), ctx.alloc(cap, np.int64)
    specs = [AggregateSpec(AggFunc.Sum, 0, 0), AggregateSpec(AggFunc.Count, 0, 1), AggregateSpec(AggFunc.Max, 1, 3, value_type=g.F64),
             AggregateSpec(AggFunc.Avg, 0, 2)]                                      # output_col_index decides the slot, not the list order
    ng = AggregateKernel(ctx).aggregate(specs, AggregateParams(dk.ptr, [di.ptr, df.ptr], n, cap, out.ptr, out_keys.ptr))
    ek, er = O.group_by_i64(k, [vi, vi, vi, vf], [O.SUM, O.COUNT, O.AVG, O.MAX])
    assert ng == len(ek)
    order = np.argsort(out_keys.to_numpy(ng))
    raw = out.to_numpy()
    slot = lambda i: raw[i * cap:i * cap + ng][order]                               # out_ptr + i * n_groups * 8 (aggregate_kernel.rs:98)
    assert (out_keys.to_numpy(ng)[order] == ek).all()
    assert (slot(0).view(np.int64) == er[0]).all() and (slot(1).view(np.int64) == er[1]).all()
    assert np.allclose(slot(2).view(np.float64), er[2], rtol=1e-12, atol=0) and (slot(3).view(np.float64) == er[3]).all()
    # no group key: the reference's global aggregate (executor.rs:307), one slot per aggregate
    one, one_key = ctx.alloc(2, np.uint64), ctx.alloc(1, np.int64)
    assert AggregateKernel(ctx).aggregate([AggregateSpec(AggFunc.Sum, 0, 0), AggregateSpec(AggFunc.Min, 0, 1)],
                                          AggregateParams(0, [di.ptr], n, 1, one.ptr, one_key.ptr)) == 1
    assert one.to_numpy().view(np.int64).tolist() == [int(vi.sum()), int(vi.min())]
