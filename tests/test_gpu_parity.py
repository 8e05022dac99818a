"""GPU parity tests: every call goes through the C ABI (include/olap_b200.h) into the sm_100a library and is
compared with the CPU oracle on the same seeded inputs.  Bit-exact for masks, selections, row-id pairs, counts
and integer/min/max aggregates; f64 SUM/AVG bit-exact on dyadic data and <= 1e-12 relative on random data."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

F64_RTOL = 1e-12  # north_star tolerance for floating-point SUM/AVG


def dev(ctx, arr):
    return ctx.upload_raw(np.ascontiguousarray(arr))


# ------------------------------------------------------------------------------------------------- generators
def test_device_generators_match_oracle(ctx):
    for kind, fn, dom in [(0, O.gen_uniform_i64, 1000), (0, O.gen_uniform_i64, 2**27), (1, O.gen_perm_i64, 100000),
                          (1, O.gen_perm_i64, 1 << 16), (2, O.gen_dyadic_f64, 1 << 20)]:
        got = ctx.generate(50001, kind, 7, dom, row0=123).to_numpy()
        assert (got == fn(50001, 7, dom, row0=123)).all()
    assert (ctx.generate(1000, 3, 9).to_numpy() == O.gen_unit_f64(1000, 9)).all()


# ------------------------------------------------------------------------------------------------- filter
def test_filter_reference_vector(ctx):
    # executor.rs:553-570
    col = dev(ctx, np.array([5, 15, 25], np.int64))
    cnt, mask, sel = ctx.filter_i64(col, 3, O.GT, 10)
    assert cnt == 2 and mask.to_numpy().tolist() == [0, 1, 1] and sel.to_numpy(cnt).tolist() == [1, 2]


@pytest.mark.parametrize("n", [0, 1, 2, 3, 2047, 2048, 2049, 100003, 1 << 20])
@pytest.mark.parametrize("op", [O.EQ, O.NE, O.LT, O.LE, O.GT, O.GE, 9])
def test_filter_i64_all_ops(ctx, n, op):
    col = O.gen_uniform_i64(n, 1, 1000) - 500
    d = dev(ctx, col)
    cnt, mask, sel = ctx.filter_i64(d, n, op, 3)
    exp = O.filter_mask_i64(col, op, literal=3)
    assert (mask.to_numpy() == exp).all()
    idx = O.mask_to_indices(exp)
    assert cnt == len(idx) and (sel.to_numpy(cnt) == idx).all()  # row order preserved


def test_filter_conjunction_columns_f64_and_unaligned(ctx):
    import gpu_olap_engine_b200 as g
    n = 300001
    a = O.gen_uniform_i64(n + 1, 2, 1 << 20); b = O.gen_uniform_i64(n + 1, 3, 1 << 20); f = O.gen_dyadic_f64(n + 1, 4, 1 << 20)
    da, db, df = dev(ctx, a), dev(ctx, b), dev(ctx, f)
    for off in (0, 1):  # off = 1: pointers are only 8-byte aligned -> scalar-load variant
        terms = [g.Term(da.ptr + 8 * off, g.GT, 1 << 19), g.Term(da.ptr + 8 * off, g.LE, rhs=db.ptr + 8 * off),
                 g.Term(df.ptr + 8 * off, g.LT, 600.25, type=g.F64)]
        cnt, mask, sel = ctx.filter(terms, n)
        aa, bb, ff = a[off:off + n], b[off:off + n], f[off:off + n]
        # conjunction = AND of the term masks (SURVEY.md 8(c) completion rule)
        exp = O.filter_mask_i64(aa, O.GT, literal=1 << 19) & O.filter_mask_i64(aa, O.LE, rhs=bb) & (ff < 600.25).astype(np.uint8)
        assert (mask.to_numpy() == exp).all() and cnt == exp.sum()
        assert (sel.to_numpy(cnt) == O.mask_to_indices(exp)).all()


def test_filter_compact_and_gather(ctx):
    import gpu_olap_engine_b200 as g
    n = 250000
    k = O.gen_uniform_i64(n, 5, 1000); v = O.gen_uniform_i64(n, 6, 1 << 20); f = O.gen_dyadic_f64(n, 7, 1 << 20)
    dk, dv, df = dev(ctx, k), dev(ctx, v), dev(ctx, f)
    exp = O.filter_mask_i64(v, O.GT, literal=1 << 19); idx = O.mask_to_indices(exp)
    cnt, outs = ctx.filter_compact([g.Term(dv, g.GT, 1 << 19)], n, [dk, dv, df], [np.int64, np.int64, np.float64])
    assert cnt == len(idx)
    for o, src in zip(outs, (k, v, f)):
        assert (o.to_numpy(cnt) == O.take(src, idx)).all()  # filter_record_batch (executor.rs:478-494)
    _, _, sel = ctx.filter_i64(dv, n, g.GT, 1 << 19, want_mask=False)
    assert (ctx.gather(df, sel, cnt, np.float64).to_numpy() == O.take(f, idx)).all()


# ------------------------------------------------------------------------------------------------- projection
@pytest.mark.parametrize("op", ["+", "-", "*", "/", "%"])
def test_project_i64(ctx, op):
    n = 100001
    a = O.gen_uniform_i64(n, 8, 2**62) - 2**61; b = O.gen_uniform_i64(n, 9, 7) - 3
    a[:4] = [-2**63, 2**63 - 1, -2**63, 5]; b[:4] = [-1, 1, 0, 0]
    da, db = dev(ctx, a), dev(ctx, b)
    assert (ctx.project_i64(da, op, n, rhs=db).to_numpy() == O.project_i64(a, op, b)).all()
    assert (ctx.project_i64(da, op, n, literal=-7).to_numpy() == O.project_i64(a, op, literal=-7)).all()


def test_project_f64_extension(ctx):
    n = 4097
    a = O.gen_dyadic_f64(n, 1, 1 << 20); b = O.gen_dyadic_f64(n, 2, 1 << 10) + 1
    da, db = dev(ctx, a), dev(ctx, b)
    assert (ctx.project_f64(da, "*", n, rhs=db).to_numpy() == a * b).all()
    assert (ctx.project_f64(da, "-", n, literal=1.0).to_numpy() == a - 1.0).all()


# ------------------------------------------------------------------------------------------------- arrow layout
def test_upload_widens_on_device(ctx):
    rng = np.random.default_rng(0)
    for dt in (np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint16, np.uint32, np.uint64, np.float32, np.float64):
        if np.issubdtype(dt, np.floating):
            a = rng.standard_normal(1001).astype(dt)
        else:
            info = np.iinfo(dt); a = rng.integers(info.min, info.max, 1001, dtype=dt, endpoint=True)
        assert (ctx.upload(a).to_numpy() == O.widen_column(a)).all()
    a = np.arange(100, dtype=np.int32); valid = (np.arange(100) % 3 != 0)
    bitmap = np.packbits(valid, bitorder="little")
    col, v = ctx.upload(a, validity_bitmap=bitmap, want_validity=True)
    assert (col.to_numpy() == O.widen_column(a, validity=valid.astype(np.uint8))).all()
    assert (v.to_numpy() == valid).all()


# ------------------------------------------------------------------------------------------------- join
def check_join(ctx, b, p, strategy, max_matches=None):
    import gpu_olap_engine_b200 as g
    db, dp = dev(ctx, b), dev(ctx, p)
    exp_c, exp_pairs = O.hash_join_i64(b, p)
    mm = max_matches if max_matches is not None else max(exp_c, 1)
    cnt, pairs = ctx.hash_join_i64(db, len(b), dp, len(p), max_matches=mm, strategy=strategy)
    assert cnt == exp_c
    got = pairs.to_numpy(2 * cnt).reshape(-1, 2)
    assert (O.sort_pairs(got) == exp_pairs).all()
    c2, _ = ctx.hash_join_i64(db, len(b), dp, len(p), strategy=strategy, want_pairs=False)
    assert c2 == exp_c


@pytest.mark.parametrize("strategy", [2, 0])  # BroadcastHashJoin, RadixHashJoin
@pytest.mark.parametrize("nb,npb,dom", [(0, 100, 10), (100, 0, 10), (1, 1, 1), (1000, 5000, 300), (5000, 1000, 100000),
                                        (4096, 4096, 17), (100000, 1000000, 100000), (300000, 70000, 50)])
def test_join_matches_oracle(ctx, strategy, nb, npb, dom):
    b = O.gen_uniform_i64(nb, 11, dom); p = O.gen_uniform_i64(npb, 12, dom)
    if nb * npb // max(dom, 1) > 40_000_000:
        pytest.skip("too many matches for a test")
    check_join(ctx, b, p, strategy)


@pytest.mark.parametrize("strategy", [2, 0])
def test_join_c1b_shape(ctx, strategy):
    # BASELINE config 1: 100 K unique build keys (permutation) joined with 1 M FK probes
    b = O.gen_perm_i64(100000, 2, 100000); p = O.gen_uniform_i64(1000000, 22, 100000)
    check_join(ctx, b, p, strategy)


@pytest.mark.parametrize("strategy", [2, 0])
def test_join_edge_keys(ctx, strategy):
    b = np.array([-1, 2**32 + 1, 0, 2**63 - 1, -2**63, 2**32, 1, 1, 1], np.int64)
    p = np.array([0, -1, 1, 2**32, 2**32 + 1, -2**63, 2**63 - 1, 99, 1], np.int64)
    check_join(ctx, b, p, strategy)
    check_join(ctx, np.full(3000, 7, np.int64), np.full(500, 7, np.int64), strategy)  # all-same key: 1.5 M pairs


@pytest.mark.parametrize("strategy", [2, 0])
def test_join_build_once_probe_in_batches(ctx, strategy):
    # the executor hands the right side over batch by batch (executor.rs:66-69): same match set as the one-shot join
    import gpu_olap_engine_b200 as g
    b = O.gen_uniform_i64(300000, 11, 200000); p = O.gen_uniform_i64(700001, 12, 250000)
    exp_c, exp_pairs = O.hash_join_i64(b, p)
    db, dp = dev(ctx, b), dev(ctx, p)
    out = ctx.alloc(2 * exp_c, np.uint32)
    total, edges = 0, [0, 100000, 100001, 433334, 700001]
    with g.JoinBuild(ctx, db, len(b), strategy) as jb:
        for lo, hi in zip(edges[:-1], edges[1:]):
            total += jb.probe(dp.ptr + 8 * lo, hi - lo, probe_row_base=lo, out_pairs=out.ptr + 8 * total, max_matches=exp_c - total)
        assert jb.probe(dp.ptr, 0) == 0
    assert total == exp_c
    assert (O.sort_pairs(out.to_numpy(2 * total).reshape(-1, 2)) == exp_pairs).all()


def test_join_overflow_is_reported(ctx):
    import gpu_olap_engine_b200 as g
    b = np.arange(1000, dtype=np.int64); p = np.arange(1000, dtype=np.int64)
    db, dp = dev(ctx, b), dev(ctx, p)
    with pytest.raises(g.OlapError) as ei:
        ctx.hash_join_i64(db, 1000, dp, 1000, max_matches=10)
    assert ei.value.status == 3 and ei.value.count == 1000  # OLAP_ERR_OVERFLOW, count still exact (join_kernels.cuh:155-160)


def test_join_large_radix_fk(ctx):
    # 4 Mi unique build keys, 16 Mi FK probes: several partitions, every probe row matches exactly once
    nb, npb = 1 << 22, 1 << 24
    db = ctx.generate(nb, 1, 3, nb); dp = ctx.generate(npb, 0, 33, nb)
    cnt, pairs = ctx.hash_join_i64(db, nb, dp, npb, max_matches=npb, strategy=0)
    assert cnt == npb
    got = pairs.to_numpy(2 * cnt).reshape(-1, 2)
    bk = db.to_numpy(); pk = dp.to_numpy()
    assert (np.sort(got[:, 1]) == np.arange(npb, dtype=np.uint32)).all()     # every probe row exactly once
    assert (bk[got[:, 0]] == pk[got[:, 1]]).all()                            # and with an equal build key
    exp_c, exp_pairs = O.hash_join_i64(bk, pk, threads=8)
    assert exp_c == cnt and (O.sort_pairs(got) == O.sort_pairs(exp_pairs)).all()


def test_join_skewed_sides_take_the_exact_layout(ctx):
    # a heavy hitter overflows its fixed-capacity partition region (one-pass layout) on either side: the join must fall back
    # to the exact histogram layout and still return the oracle's match set
    rng = np.random.default_rng(5)
    b = rng.integers(0, 1 << 40, 200000).astype(np.int64); b[: 160000] = 424242           # 80 % of the build side is one key
    p = rng.integers(0, 1 << 40, 300000).astype(np.int64); p[:3] = 424242; p[1000:271000] = b[199999]  # 90 % of the probe side too
    rng.shuffle(b); rng.shuffle(p)
    check_join(ctx, b, p, 0)


@pytest.mark.timeout(180)
def test_join_large_skewed_partition_keeps_three_ring_slots(ctx):
    # 4 partitions, one of them > 1 Mi build rows (a 40 MB table): the pipeline's table ring must not shrink below 3 slots
    # (with 2 the clear of partition p would wait for probe tiles queued behind it -- a deadlock); unique heavy-partition keys
    # are produced by giving half of the build side keys that all hash to the same radix partition
    import gpu_olap_engine_b200 as g
    lib = g.load()
    rng = np.random.default_rng(9)
    cand = rng.integers(0, 1 << 50, 6_000_000).astype(np.int64)
    part = np.array([lib.olap_partition_of(int(k), 2) for k in cand[:200]])   # spot check of the vectorised formula below
    u = cand.astype(np.uint64); h = ((u ^ (u >> np.uint64(29))) * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(62)
    assert (h[:200] == part).all()
    heavy = cand[h == 0][:1_200_000]; rest = cand[h != 0][:800_000]
    b = np.concatenate([heavy, rest]); rng.shuffle(b)
    p = np.concatenate([b[:500_000], rng.integers(0, 1 << 50, 300_000).astype(np.int64)])
    db, dp = dev(ctx, b), dev(ctx, p)
    exp_c, exp_pairs = O.hash_join_i64(b, p, threads=8)
    cnt, pairs = ctx.hash_join_i64(db, len(b), dp, len(p), max_matches=exp_c, strategy=0)
    assert cnt == exp_c
    assert (O.sort_pairs(pairs.to_numpy(2 * cnt).reshape(-1, 2)) == O.sort_pairs(exp_pairs)).all()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("part_rows", ["524288", "1024"])   # 2 partitions / up to 256 tiny partitions (ring reuse, dependency counters)
@pytest.mark.parametrize("alloc", ["1", "0"])   # one-pass fixed-capacity regions / exact histogram layout (unaligned partition starts)
def test_join_tile_boundaries_in_both_partition_layouts(ctx, alloc, part_rows, monkeypatch):
    # radix strategy on sizes that straddle the 2048-row work items, the 4096-row partition tiles and the 4-row vector
    # loads; every size is checked against the oracle's match set
    monkeypatch.setenv("OLAP_JOIN_ALLOC", alloc)
    monkeypatch.setenv("OLAP_JOIN_PART_ROWS", part_rows)
    rng = np.random.default_rng(int(alloc) + 17)
    for nb, npb in [(2048, 4096), (2049, 4095), (4097, 8191), (8190, 3), (3, 8190), (65537, 131071), (131073, 70001), (1 << 18, (1 << 19) + 5)]:
        dom = max(nb // 2, 2)
        b = rng.integers(-dom, dom, nb).astype(np.int64); p = rng.integers(-dom, dom, npb).astype(np.int64)
        check_join(ctx, b, p, 0)


@pytest.mark.timeout(300)
def test_join_direct_address_tables_and_their_fallbacks(ctx):
    # >= 2^20 build rows with a dense key range take the direct-address path (row_id[key - min] per key range); a repeated
    # build key, or a probe side whose key ranges are very uneven, sends the join back to the bucket tables.  Same match set.
    rng = np.random.default_rng(23)
    nb, npb = (1 << 20) + 321, (1 << 21) + 77
    perm = rng.permutation(nb).astype(np.int64)
    cases = {
        "dense unique, negative minimum": (perm - 500_000, rng.integers(-600_000, nb - 400_000, npb).astype(np.int64)),
        "unique with holes (range = 3 x rows), probes in the holes and outside": (perm * 3 + 7, rng.integers(-50, 3 * nb + 100, npb).astype(np.int64)),
        "a quarter of a dense domain (range = 4 x rows), as a rank of a 4-GPU join sees it": (perm * 4 + 1, rng.integers(0, 4 * nb + 8, npb).astype(np.int64)),
        "dense with duplicates -> bucket tables": (rng.integers(0, nb, nb).astype(np.int64), rng.integers(0, nb, npb // 2).astype(np.int64)),
        "one hot probe key -> uneven key ranges -> bucket tables": (perm, np.where(rng.random(npb) < 0.7, 12345, rng.integers(0, nb, npb)).astype(np.int64)),
    }
    for name, (b, p) in cases.items():
        db, dp = dev(ctx, b), dev(ctx, p)
        exp_c, exp_pairs = O.hash_join_i64(b, p, threads=8)
        cnt, pairs = ctx.hash_join_i64(db, len(b), dp, len(p), max_matches=max(exp_c, 1), strategy=0)
        assert cnt == exp_c, name
        assert (O.sort_pairs(pairs.to_numpy(2 * cnt).reshape(-1, 2)) == O.sort_pairs(exp_pairs)).all(), name


@pytest.mark.parametrize("bits", [1, 4, 8])
def test_radix_partition_unit(ctx, bits):
    import gpu_olap_engine_b200 as g
    n = 300007
    keys = O.gen_uniform_i64(n, 13, 2**40) - 2**39
    dk = dev(ctx, keys)
    ok, oi, begin = ctx.radix_partition_i64(dk, n, bits)
    ok, oi, begin = ok.to_numpy(), oi.to_numpy(), begin.to_numpy()
    P = 1 << bits
    assert begin[0] == 0 and begin[P] == n and (np.diff(begin.astype(np.int64)) >= 0).all()
    assert (keys[oi] == ok).all() and (np.sort(oi) == np.arange(n)).all()     # a permutation of the input rows
    part = np.array([g.load().olap_partition_of(int(k), bits) for k in ok[::997]])
    pos = np.arange(n)[::997]
    assert ((begin[part] <= pos) & (pos < begin[part + 1])).all()           # every row sits in its hash partition


# ------------------------------------------------------------------------------------------------- group-by
def check_group_by(ctx, keys, cols, funcs, hint, terms_fn=None, mask=None, rtol=0.0):
    import gpu_olap_engine_b200 as g
    n = len(cols[0]) if cols else len(keys)
    dk = None if keys is None else dev(ctx, keys)
    uploaded = {}
    dcols = [uploaded.setdefault(id(c), dev(ctx, c)) if id(c) not in uploaded else uploaded[id(c)] for c in cols]   # one device column per distinct input
    aggs = [(f, d, g.F64 if c.dtype == np.float64 else g.I64) for f, d, c in zip(funcs, dcols, cols)]
    terms = terms_fn(dcols) if terms_fn else None
    ng, ok, oa, cap = ctx.group_by_i64(dk, n, aggs, hint, out_capacity=max(hint, 1 << 16), terms=terms)
    gk = ok.to_numpy(ng); res = g.typed_agg_results(oa.to_numpy(), cap, ng, aggs)
    order = np.argsort(gk, kind="stable")
    if mask is not None:
        sel = mask.astype(bool)
        ek, eres = O.group_by_i64(None if keys is None else keys[sel], [c[sel] for c in cols], funcs)
    else:
        ek, eres = O.group_by_i64(keys, cols, funcs)
    assert ng == len(ek) and (gk[order] == ek).all()
    for got, exp, f, c in zip(res, eres, funcs, cols):
        got = got[order]
        if got.dtype == np.float64 and rtol:
            assert np.allclose(got, exp, rtol=rtol, atol=0), f
        else:
            assert (got == exp).all(), (f, c.dtype)
    return ng


ALL5 = [O.SUM, O.COUNT, O.MIN, O.MAX, O.AVG]


@pytest.mark.parametrize("G,hint", [(1, 1), (7, 7), (100, 100), (1000, 1000), (3000, 3000), (1000, 10), (50000, 50000), (50000, 64)])
def test_group_by_i64_regimes(ctx, G, hint):
    # hint == G: the intended regime; hint << G: the hint is wrong and correctness must not depend on it
    n = 400003
    k = O.gen_uniform_i64(n, 4, G) * 7919 - 1000; v = O.gen_uniform_i64(n, 5, 1 << 20) - (1 << 19)
    check_group_by(ctx, k, [v] * 5, ALL5, hint)


@pytest.mark.parametrize("G,hint", [(50000, 50000), (50000, 20000), (50000, 3000), (700000, 700000), (3000000, 3000000)])
def test_group_by_partitioned_regime(ctx, G, hint):
    # >= 2^20 rows and more groups than a shared-memory table holds: radix-partitioned aggregation.  hint < G makes the
    # per-partition tables overflow (rows spill to the global table) -- the result must not change
    n = (1 << 21) + 5
    k = O.gen_uniform_i64(n, 4, G) * 7919 - 1000; v = O.gen_uniform_i64(n, 5, 1 << 20) - (1 << 19); f = O.gen_dyadic_f64(n, 6, 1 << 20)
    check_group_by(ctx, k, [v, v, v, v, v, f, f], ALL5 + [O.SUM, O.MIN], hint)


def test_group_by_partitioned_skewed_keys_take_the_exact_layout(ctx):
    # 60 % of the rows share one key: its fixed-capacity partition region overflows, the partition pass must fall back to
    # the exact histogram layout and the aggregates must not change
    n = (1 << 21) + 77
    k = O.gen_uniform_i64(n, 4, 50000) * 31 + 5; v = O.gen_uniform_i64(n, 5, 1 << 20) - (1 << 19)
    k[: int(n * 0.6)] = 123456789
    np.random.default_rng(3).shuffle(k)
    check_group_by(ctx, k, [v, v, v, v], [O.SUM, O.COUNT, O.MIN, O.MAX], 50000)


def test_group_by_partitioned_with_fused_filter(ctx):
    import gpu_olap_engine_b200 as g
    n = 1 << 21
    k = O.gen_uniform_i64(n, 4, 40000); v = O.gen_uniform_i64(n, 5, 1 << 20)
    mask = O.filter_mask_i64(v, O.GT, literal=1 << 19)
    check_group_by(ctx, k, [v, v], [O.SUM, O.COUNT], 40000, terms_fn=lambda d: [g.Term(d[0], g.GT, 1 << 19)], mask=mask)


def test_group_by_f64_dyadic_bit_exact_and_random_tolerance(ctx):
    n = 300000
    k = O.gen_uniform_i64(n, 4, 500); f = O.gen_dyadic_f64(n, 6, 1 << 20); r = O.gen_unit_f64(n, 7) - 0.5
    check_group_by(ctx, k, [f] * 5, ALL5, 500)                               # bit-exact: sums are order-independent
    import gpu_olap_engine_b200 as g
    dk, dr = dev(ctx, k), dev(ctx, r)
    aggs = [(g.SUM, dr, g.F64), (g.AVG, dr, g.F64), (g.MIN, dr, g.F64), (g.MAX, dr, g.F64)]
    ng, ok, oa, cap = ctx.group_by_i64(dk, n, aggs, 500)
    gk = ok.to_numpy(ng); res = g.typed_agg_results(oa.to_numpy(), cap, ng, aggs); order = np.argsort(gk)
    ek, eres = O.group_by_i64(k, [r] * 4, [O.SUM, O.AVG, O.MIN, O.MAX], compensated=True)
    scale = np.array([np.abs(r[k == kk]).sum() for kk in ek])               # relative to the magnitude summed
    assert (np.abs(res[0][order] - eres[0]) <= F64_RTOL * scale).all()
    assert (np.abs(res[1][order] - eres[1]) <= F64_RTOL * scale).all()
    assert (res[2][order] == eres[2]).all() and (res[3][order] == eres[3]).all()


def test_group_by_global_matches_reference_execution(ctx):
    # what execute_aggregate actually computes today (executor.rs:307-346): one global group
    n = 1 << 20
    v = O.gen_uniform_i64(n, 5, 1 << 20); f = O.gen_dyadic_f64(n, 6, 1 << 20)
    check_group_by(ctx, None, [v] * 5 + [f] * 5, ALL5 + ALL5, 1)
    import gpu_olap_engine_b200 as g
    dv = dev(ctx, v)
    aggs = [(fn, dv, g.I64) for fn in ALL5]
    ng, ok, oa, cap = ctx.group_by_i64(None, n, aggs, 1)
    res = g.typed_agg_results(oa.to_numpy(), cap, ng, aggs)
    ref = O.GlobalAgg(); ref.update(v)
    for got, fn in zip(res, ALL5):
        assert float(got[0]) == ref.final(fn)      # reference-compat view: identical while magnitudes < 2^53
    # empty input: SUM 0, COUNT 0, AVG 0.0, MIN/MAX identities (executor.rs:310-311, :339-345; aggregate_kernel.rs:27-33)
    ng, ok, oa, cap = ctx.group_by_i64(None, 0, aggs, 1)
    res = g.typed_agg_results(oa.to_numpy(), cap, ng, aggs)
    assert ng == 1 and [int(r[0]) if r.dtype != np.float64 else float(r[0]) for r in res] == [0, 0, 2**63 - 1, -2**63, 0.0]


def test_group_by_fused_filter_c1a_shape(ctx):
    # BASELINE config 1: filter v_i > 2^19 then GROUP BY k SUM/COUNT on a 1 Mi-row table
    import gpu_olap_engine_b200 as g
    n = 1 << 20
    k = O.gen_uniform_i64(n, 1, 1000); vi = O.gen_uniform_i64(n, 11, 1 << 20); vf = O.gen_dyadic_f64(n, 12, 1 << 20)
    mask = O.filter_mask_i64(vi, O.GT, literal=1 << 19)
    check_group_by(ctx, k, [vi, vi, vf], [O.SUM, O.COUNT, O.SUM], 1000, terms_fn=lambda d: [g.Term(d[0], g.GT, 1 << 19)], mask=mask)
    check_group_by(ctx, None, [vi, vi, vf], [O.SUM, O.COUNT, O.AVG], 1, terms_fn=lambda d: [g.Term(d[0], g.GT, 1 << 19)], mask=mask)


def test_group_by_edge_keys_and_empty(ctx):
    k = np.array([-2**63, 2**63 - 1, -2**63, 0, -1, 0, 2**63 - 1, -2**63], np.int64)
    v = np.arange(8, dtype=np.int64) - 3
    check_group_by(ctx, k, [v] * 5, ALL5, 4)
    check_group_by(ctx, k, [v] * 5, ALL5, 100000)                            # forces the global-table regime
    assert check_group_by(ctx, np.array([], np.int64), [np.array([], np.int64)], [O.SUM], 10) == 0


def test_group_by_capacity_error(ctx):
    import gpu_olap_engine_b200 as g
    k = dev(ctx, np.arange(1000, dtype=np.int64))
    with pytest.raises(g.OlapError) as ei:
        ctx.group_by_i64(k, 1000, [(g.COUNT, None, g.I64)], 1000, out_capacity=10)
    assert ei.value.status == 6


# ------------------------------------------------------------------------------------------------- memory
def test_pool_alloc_free_and_cap():
    import gpu_olap_engine_b200 as g
    with g.Context(0, max_gpu_memory=1 << 20) as c:   # EngineConfig.max_gpu_memory as the pool cap
        a = c.alloc(1000, np.int64)
        assert c.pool_stats()["used"] >= 8000
        with pytest.raises(g.OlapError) as ei:
            c.alloc(1 << 20, np.int64)
        assert ei.value.status == 2
        a.free()
        assert c.pool_stats()["used"] == 0 and c.pool_stats()["high_water"] >= 8000


def test_group_by_expression_programs_at_the_kernel_seam(ctx):
    # olap_group_by_prog_i64: SUM / MIN / MAX over a*3 - a/b (i64: wrapping, truncating division, x/0 = 0) and f*(0.5+f) (f64),
    # evaluated inside the aggregation kernel, against the oracle's projection + group-by; shared, no-key and global regimes
    import gpu_olap_engine_b200 as g
    n = 400003
    k = O.gen_uniform_i64(n, 80, 700) - 350
    a = O.gen_uniform_i64(n, 81, 1 << 30) - (1 << 29); b = O.gen_uniform_i64(n, 82, 7) - 3   # b hits 0: x / 0 = 0
    f = O.gen_dyadic_f64(n, 83, 1 << 12)
    e1 = O.project_i64(O.project_i64(a, "*", literal=3), "-", r=O.project_i64(a, "/", r=b))
    e2 = f * (0.5 + f)
    dk, da, db, df = dev(ctx, k), dev(ctx, a), dev(ctx, b), dev(ctx, f)
    p1 = [("c", da), ("l", 3), "*", ("c", da), ("c", db), "/", "-"]
    p2 = [("c", df), ("l", 0.5), ("c", df), "+", "*"]
    aggs = [(g.SUM, None, g.I64), (g.MIN, None, g.I64), (g.MAX, None, g.I64), (g.SUM, None, g.F64), (g.COUNT, None, g.I64)]
    progs = [p1, p1, p1, p2, None]
    for keys, hint in ((dk, 700), (None, 1), (dk, 70000)):
        exp = O.group_by_i64(k if keys is not None else None, [e1, e1, e1, e2, e1], [O.SUM, O.MIN, O.MAX, O.SUM, O.COUNT])
        ng, ok, oa, cap = ctx.group_by_i64(keys, n, aggs, hint, progs=progs)
        assert ng == len(exp[0])
        gk = ok.to_numpy(ng); order = np.argsort(gk, kind="stable")
        res = g.typed_agg_results(oa.to_numpy(), cap, ng, aggs)
        eo = np.argsort(exp[0], kind="stable")
        if keys is not None:
            assert (gk[order] == np.asarray(exp[0])[eo]).all()
        for got, want in zip(res, exp[1]):
            assert (np.asarray(got)[order] == np.asarray(want)[eo]).all()



def test_group_by_computed_key_at_the_kernel_seam(ctx):
    # olap_group_by_expr_i64: GROUP BY a * 3 + b evaluated inside the aggregation kernel (no key column) == the oracle's group-by over the
    # materialised key; shared regime, global regime (many groups, few rows), with a fused filter and a value program next to it
    import gpu_olap_engine_b200 as g
    for n, da_, db_, hint in ((500_003, 40, 3, 200), (60_001, 30_000, 3, 100_000)):
        a = O.gen_uniform_i64(n, 90, da_) - da_ // 2; b = O.gen_uniform_i64(n, 91, db_)
        v = O.gen_uniform_i64(n, 92, 1 << 20); f = O.gen_dyadic_f64(n, 93, 1 << 12)
        key = O.project_i64(O.project_i64(a, "*", literal=3), "+", r=b)
        dA, dB, dV, dF = dev(ctx, a), dev(ctx, b), dev(ctx, v), dev(ctx, f)
        kprog = [("c", dA), ("l", 3), "*", ("c", dB), "+"]
        aggs = [(g.SUM, dV, g.I64), (g.COUNT, None, g.I64), (g.MAX, None, g.F64), (g.AVG, dV, g.I64)]
        progs = [None, None, [("c", dF), ("l", 2.0), "*"], None]
        keep = v > (1 << 19)
        for terms, mask in ((None, np.ones(n, bool)), ([g.Term(dV, g.GT, 1 << 19)], keep)):
            ek, er = O.group_by_i64(key[mask], [v[mask], v[mask], (f * 2.0)[mask], v[mask]], [O.SUM, O.COUNT, O.MAX, O.AVG])
            ng, ok, oa, cap = ctx.group_by_i64(None, n, aggs, hint, terms=terms, progs=progs, key_prog=kprog)
            assert ng == len(ek)
            gk = ok.to_numpy(ng); order = np.argsort(gk)
            assert (gk[order] == ek).all()
            for i, (got, want) in enumerate(zip(g.typed_agg_results(oa.to_numpy(), cap, ng, aggs), er)):
                assert np.allclose(np.asarray(got)[order], want, rtol=1e-12, atol=0) if i == 3 else (np.asarray(got)[order] == want).all(), i
    # a cardinality that needs the partitioned regime is refused (the caller materialises the key), as for value programs
    n = 1 << 21
    a = O.gen_uniform_i64(n, 94, 1 << 40); dA = dev(ctx, a)
    with pytest.raises(g.OlapError) as ei:
        ctx.group_by_i64(None, n, [(g.COUNT, None, g.I64)], n, key_prog=[("c", dA), ("l", 1), "+"])
    assert ei.value.status == 5
