"""GPU parity of the direct-address join with shared-memory tables (csrc/djoin.cu) and of the paths around it: the
remembered key range, heavy-hitter handling for Zipf-skewed probe keys, the build-once / probe-per-batch API on dense keys,
and the table-ring ordering of the bucket-table pipeline (round-1 advisor finding).  Oracle: oracle/oracle.c
(executor.rs:238-253 + probe_hash_table_kernel semantics, join_kernels.cuh:146-165); match sets compared as sorted pairs."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


def dev(ctx, a):
    return ctx.upload_raw(np.ascontiguousarray(a))


def join_and_check(ctx, b, p, name="", strategy=0, db=None, dp=None):
    ctx.set_profiling(True)
    db = dev(ctx, b) if db is None else db
    dp = dev(ctx, p) if dp is None else dp
    exp_c, exp_pairs = O.hash_join_i64(b, p, threads=8)
    cnt, pairs = ctx.hash_join_i64(db, len(b), dp, len(p), max_matches=max(exp_c, 1), strategy=strategy)
    assert cnt == exp_c, name
    assert (O.sort_pairs(pairs.to_numpy(2 * cnt).reshape(-1, 2)) == O.sort_pairs(exp_pairs)).all(), name
    return ctx.last_phase_times()


def zipf_keys(rng, n, domain, s, perm_seed=7):
    """n draws from Zipf(s) over `domain` ranks, ranks mapped to keys through a fixed permutation."""
    ranks = np.arange(1, domain + 1, dtype=np.float64)
    w = ranks ** (-s)
    cdf = np.cumsum(w); cdf /= cdf[-1]
    r = np.searchsorted(cdf, rng.random(n))
    perm = np.random.default_rng(perm_seed).permutation(domain)
    return perm[np.minimum(r, domain - 1)].astype(np.int64)


@pytest.mark.timeout(300)
def test_smem_direct_join_is_the_path_taken_and_matches(ctx):
    rng = np.random.default_rng(31)
    nb, npb = (1 << 21) + 13, (1 << 22) + 5
    perm = rng.permutation(nb).astype(np.int64)
    phases = join_and_check(ctx, perm - 77, rng.integers(-1000, nb + 1000, npb).astype(np.int64), "dense")
    assert "build_probe_smem" in phases, phases
    # tile / vector boundaries of the partition pass and the probe groups
    for nb2, np2 in [(1 << 20, 1), (1 << 20, 8191), ((1 << 20) + 1, 8193), ((1 << 20) + 4097, 2049)]:
        b = rng.permutation(nb2).astype(np.int64) + 5
        join_and_check(ctx, b, rng.integers(0, nb2 + 10, np2).astype(np.int64), f"{nb2} x {np2}")


@pytest.mark.timeout(300)
def test_smem_direct_join_truncates_at_max_matches_with_exact_count(ctx):
    import gpu_olap_engine_b200 as g
    rng = np.random.default_rng(32)
    nb, npb = 1 << 20, 1 << 21
    b = rng.permutation(nb).astype(np.int64); p = rng.integers(0, nb, npb).astype(np.int64)
    db, dp = dev(ctx, b), dev(ctx, p)
    with pytest.raises(g.OlapError) as ei:
        ctx.hash_join_i64(db, nb, dp, npb, max_matches=1000, strategy=0)
    assert ei.value.status == 3 and ei.value.count == npb
    c, _ = ctx.hash_join_i64(db, nb, dp, npb, strategy=0, want_pairs=False)
    assert c == npb


@pytest.mark.timeout(300)
def test_remembered_key_range_is_verified_by_the_build_pass(ctx):
    # the min / max of a build column is remembered per (pointer, rows); new contents at the same address with keys outside the
    # remembered range must be noticed (flag raised by the partition pass) and the join redone with a fresh range
    rng = np.random.default_rng(33)
    nb, npb = 1 << 20, 1 << 20
    b1 = rng.permutation(nb).astype(np.int64)
    db = dev(ctx, b1)
    p1 = rng.integers(0, nb, npb).astype(np.int64)
    ph = join_and_check(ctx, b1, p1, "first contents", db=db)
    assert "minmax" in ph
    ph = join_and_check(ctx, b1, p1, "same contents again", db=db)
    assert "minmax" not in ph and "build_probe_smem" in ph          # no second min / max pass
    b2 = b1 + 3_000_000                                             # same buffer, every key outside the remembered range
    ctx.lib.olap_memcpy_h2d(ctx.handle, db.ptr, b2.ctypes.data, nb * 8, None); ctx.synchronize()
    p2 = rng.integers(2_900_000, 3_000_000 + nb + 1000, npb).astype(np.int64)
    ph = join_and_check(ctx, b2, p2, "new contents, stale range", db=db)
    assert "minmax" in ph
    b3 = b2.copy(); b3[::2] -= 1_000_000                            # wider than remembered, half inside
    ctx.lib.olap_memcpy_h2d(ctx.handle, db.ptr, b3.ctypes.data, nb * 8, None); ctx.synchronize()
    join_and_check(ctx, b3, p2, "partly outside", db=db)


@pytest.mark.timeout(600)
@pytest.mark.parametrize("s", [0.5, 1.0, 1.5])
def test_zipf_skewed_probe_keys(ctx, s):
    # BASELINE config 5's skew on one GPU: heavy hitters are sampled, resolved to their build rows and joined in the partition
    # pass itself; the rest goes through the key-range partitions
    rng = np.random.default_rng(40 + int(s * 10))
    nb, npb = 1 << 21, 1 << 23
    b = rng.permutation(nb).astype(np.int64)
    p = zipf_keys(rng, npb, nb, s)
    ph = join_and_check(ctx, b, p, f"zipf {s}")
    if s >= 1.0:
        assert "hot_keys" in ph and "build_probe_smem" in ph, ph


@pytest.mark.timeout(300)
def test_hot_probe_keys_without_a_build_row_and_at_the_range_ends(ctx):
    rng = np.random.default_rng(41)
    nb, npb = 1 << 20, 1 << 22
    b = rng.permutation(nb).astype(np.int64) * 2          # even keys only: odd keys are holes
    p = rng.integers(0, 2 * nb, npb).astype(np.int64)
    hot = rng.random(npb)
    p[hot < 0.20] = 0                                     # smallest key, present
    p[(hot >= 0.20) & (hot < 0.40)] = 2 * (nb - 1)        # largest key, present
    p[(hot >= 0.40) & (hot < 0.55)] = 12345               # odd: a heavy hitter with no build row
    p[(hot >= 0.55) & (hot < 0.60)] = 2 * nb + 999        # outside the build range
    join_and_check(ctx, b, p, "hot keys with and without build rows")


@pytest.mark.timeout(600)
def test_more_than_4096_key_ranges_use_the_wide_partition_kernel(ctx):
    # range = 5 x 2^25 keys -> 5120 ranges of 2^15 keys: the 1024-thread instantiation of the partition pass
    rng = np.random.default_rng(42)
    nb, npb = 1 << 25, 1 << 25
    b = rng.permutation(nb).astype(np.int64) * 5 - 1000
    p = rng.integers(-2000, 5 * nb, npb).astype(np.int64)
    ph = join_and_check(ctx, b, p, "5120 ranges")
    assert "build_probe_smem" in ph


@pytest.mark.timeout(300)
def test_build_once_probe_in_batches_on_dense_keys(ctx):
    import gpu_olap_engine_b200 as g
    rng = np.random.default_rng(43)
    nb, npb = 1 << 20, (1 << 22) + 3
    b = rng.permutation(nb).astype(np.int64) + 10
    p = rng.integers(0, nb + 20, npb).astype(np.int64)
    exp_c, exp_pairs = O.hash_join_i64(b, p, threads=8)
    db, dp = dev(ctx, b), dev(ctx, p)
    out = ctx.alloc(2 * exp_c, np.uint32)
    total, edges = 0, [0, 1 << 20, (1 << 20) + 1, 3_000_000, npb]
    with g.JoinBuild(ctx, db, nb, 0) as jb:
        for lo, hi in zip(edges[:-1], edges[1:]):
            total += jb.probe(dp.ptr + 8 * lo, hi - lo, probe_row_base=lo, out_pairs=out.ptr + 8 * total, max_matches=exp_c - total)
    assert total == exp_c
    assert (O.sort_pairs(out.to_numpy(2 * total).reshape(-1, 2)) == O.sort_pairs(exp_pairs)).all()
    # a repeated build key: the handle falls back to bucket tables on the first probe
    b2 = b.copy(); b2[5] = b2[6]
    db2 = dev(ctx, b2)
    exp_c, exp_pairs = O.hash_join_i64(b2, p, threads=8)
    out = ctx.alloc(2 * exp_c, np.uint32)
    with g.JoinBuild(ctx, db2, nb, 0) as jb:
        c = jb.probe(dp.ptr, npb, out_pairs=out.ptr, max_matches=exp_c)
    assert c == exp_c and (O.sort_pairs(out.to_numpy(2 * c).reshape(-1, 2)) == O.sort_pairs(exp_pairs)).all()


# ---- table-ring ordering of the bucket / L2 direct pipeline: an inactive partition between two active ones
@pytest.mark.timeout(600)
def test_pipeline_ring_slot_reuse_skips_inactive_partitions_bucket_tables(ctx, monkeypatch):
    # >= 4 Mi build rows -> >= 8 hash partitions on a ring of 4 slots; probe keys confined to the partitions 1, 5 (share a slot;
    # 9 would too) so that partitions in between are inactive and clear(5) has nothing to wait on but probe(1)
    import gpu_olap_engine_b200 as g
    monkeypatch.setenv("OLAP_JOIN_DIRECT", "0")
    lib = g.load()
    rng = np.random.default_rng(44)
    nb = 1 << 23
    b = rng.integers(0, 1 << 50, nb).astype(np.int64)
    bits = 4  # 2^23 rows / 512 Ki rows per partition = 16 partitions
    u = b.astype(np.uint64); h = ((u ^ (u >> np.uint64(29))) * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(64 - bits)
    assert all(int(h[i]) == lib.olap_partition_of(int(b[i]), bits) for i in range(0, 4000, 97))
    for parts in [(1, 5), (2, 10), (0, 4, 12), (3,)]:
        cand = b[np.isin(h, parts)]
        p = np.concatenate([cand[rng.integers(0, len(cand), 600_000)], rng.integers(0, 1 << 50, 1000).astype(np.int64)])
        p = p[np.isin(((p.astype(np.uint64) ^ (p.astype(np.uint64) >> np.uint64(29))) * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(64 - bits), parts)]
        join_and_check(ctx, b, p, f"probe keys only in partitions {parts}")


@pytest.mark.timeout(600)
def test_pipeline_ring_slot_reuse_skips_inactive_partitions_l2_direct_tables(ctx, monkeypatch):
    # L2-resident direct tables (shared-memory path switched off): two dense clusters with a hole of >= 2^16 keys between them
    # and probe keys confined to a few key ranges
    monkeypatch.setenv("OLAP_JOIN_SMEM", "0")
    rng = np.random.default_rng(45)
    nb = 1 << 21
    half = nb // 2
    b = np.concatenate([rng.permutation(half), rng.permutation(half) + half + 6 * (1 << 16)]).astype(np.int64)
    rng.shuffle(b)
    ranges = [(0, 1 << 16), (3 << 16, 4 << 16), (half + 6 * (1 << 16), half + 7 * (1 << 16)), (half + 9 * (1 << 16), half + 10 * (1 << 16))]
    p = np.concatenate([rng.integers(lo, hi, 300_000) for lo, hi in ranges]).astype(np.int64)
    rng.shuffle(p)
    join_and_check(ctx, b, p, "clusters with a hole")
    # and with probe keys over the whole domain, so that the L2 direct tables are the path taken
    ph = join_and_check(ctx, b, rng.integers(-5, nb + 7 * (1 << 16), 1 << 21).astype(np.int64), "clusters with a hole, probes everywhere")
    assert "build_probe_direct" in ph, ph
