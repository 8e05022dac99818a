"""The range-sharded join's exchange on ONE GPU: all ranks of the group live in one process, each with its own context and
streams on device 0 (olap_shard_ctx_create_all with repeated devices), ordered by stream events -- the same kernels, slab
layout and phase order as the one-process-per-GPU path (tests/run_sharded_nccl.py), so the driver's 1-GPU box covers the
exchange logic.  Oracle: the global join on the CPU (oracle/oracle.c), pairs compared as sorted global row ids."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


def run_group(world, chunks, build_shards, probe_shards, expect_unsupported=False):
    import gpu_olap_engine_b200 as g
    from gpu_olap_engine_b200.shard import ShardGroup
    nb = [len(b) for b in build_shards]; npb = [len(p) for p in probe_shards]
    gb, gp = np.concatenate(build_shards), np.concatenate(probe_shards)
    exp_c, exp_pairs = O.hash_join_i64(gb, gp, threads=8)
    with ShardGroup([0] * world, max(nb), max(max(npb), 1), chunks) as grp:
        db = [grp.ctxs[r].upload_raw(build_shards[r]) for r in range(world)]
        dp = [grp.ctxs[r].upload_raw(probe_shards[r]) for r in range(world)]
        mm = [max(2 * sum(npb) // world + 4096, 4096)] * world
        out = [grp.ctxs[r].alloc(2 * mm[r], np.uint32) for r in range(world)]
        if expect_unsupported:
            with pytest.raises(g.OlapError) as ei:
                grp.join(db, nb, dp, npb, out, mm)
            assert ei.value.status == 5
            return
        for rep in range(2):   # twice: the second join reuses the slabs and the remembered key ranges
            counts = grp.join(db, nb, dp, npb, out, mm)
            assert sum(counts) == exp_c, (counts, exp_c)
            boff = np.concatenate([[0], np.cumsum(nb)[:-1]]); poff = np.concatenate([[0], np.cumsum(npb)[:-1]])
            got = []
            for r in range(world):
                left, right = grp.pairs_to_global(r, out[r], counts[r], boff, poff)
                got.append(np.stack([left, right], 1))
            got = np.concatenate(got).astype(np.uint32)
            assert (O.sort_pairs(got) == O.sort_pairs(exp_pairs)).all()
        return counts


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world,chunks", [(2, 1), (2, 3), (4, 2), (8, 2)])
def test_range_sharded_join_in_process(world, chunks):
    rng = np.random.default_rng(100 + world + chunks)
    nb, npb = (1 << 18) + 17 * world, (1 << 20) + 5
    perm = rng.permutation(nb * world).astype(np.int64) - 12345          # dense global domain with a negative minimum
    builds = [perm[r * nb:(r + 1) * nb] for r in range(world)]
    probes = [rng.integers(-12345 - 50, nb * world - 12345 + 50, npb).astype(np.int64) for _ in range(world)]   # a few keys outside the domain
    run_group(world, chunks, builds, probes)


@pytest.mark.timeout(600)
def test_range_sharded_join_uneven_shards_holes_and_empty_ranks():
    rng = np.random.default_rng(7)
    world = 4
    keys = rng.permutation(3 * (1 << 19)).astype(np.int64)[: 1 << 19] * 1 + 1000    # unique, range = 3 x rows (holes)
    cut = [0, 100_000, 100_000, 400_000, 1 << 19]                                   # rank 1 holds no build rows
    builds = [keys[cut[r]:cut[r + 1]] for r in range(world)]
    probes = [rng.integers(0, 3 * (1 << 19) + 2000, n).astype(np.int64) for n in (700_001, 0, 1_200_000, 3)]   # rank 1 holds no probe rows either
    run_group(world, 2, builds, probes)


@pytest.mark.timeout(600)
def test_range_sharded_join_reports_what_it_cannot_serve():
    rng = np.random.default_rng(8)
    world = 2
    sparse = [rng.integers(0, 1 << 50, 1 << 18).astype(np.int64) for _ in range(world)]
    probes = [rng.integers(0, 1 << 50, 1 << 18).astype(np.int64) for _ in range(world)]
    run_group(world, 1, sparse, probes, expect_unsupported=True)             # not a dense domain
    dup = [np.arange(1 << 18, dtype=np.int64), np.arange(1 << 18, dtype=np.int64) + (1 << 18)]
    dup[1][5] = dup[1][6]                                                    # a repeated build key: the shared-memory tables cannot hold it
    run_group(world, 1, dup, [rng.integers(0, 1 << 19, 1 << 18).astype(np.int64) for _ in range(world)], expect_unsupported=True)


def zipf_keys(rng, n, keys_by_rank, s):
    """n draws: Zipf(s) ranks mapped to keys through the (pseudo-random) order of keys_by_rank."""
    dom = len(keys_by_rank)
    w = np.arange(1, dom + 1, dtype=np.float64) ** (-s)
    cdf = np.cumsum(w); cdf /= cdf[-1]
    return keys_by_rank[np.minimum(np.searchsorted(cdf, rng.random(n)), dom - 1)]


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world,s", [(2, 1.0), (4, 1.5), (4, 0.5), (8, 1.0)])
def test_range_sharded_join_zipf_heavy_hitters_stay_on_their_rank(world, s):
    import gpu_olap_engine_b200 as g
    from gpu_olap_engine_b200.shard import ShardGroup
    rng = np.random.default_rng(300 + world)
    nb, npb = 1 << 18, 1 << 20
    perm = rng.permutation(nb * world).astype(np.int64) + 5
    builds = [perm[r * nb:(r + 1) * nb] for r in range(world)]
    probes = [zipf_keys(rng, npb, perm, s) for _ in range(world)]
    probes[0][:1000] = 3          # a key below the domain's hot keys, absent from the build side (min is 5)
    gb, gp = np.concatenate(builds), np.concatenate(probes)
    exp_c, exp_pairs = O.hash_join_i64(gb, gp, threads=8)
    with ShardGroup([0] * world, nb, npb, 2) as grp:
        db = [grp.ctxs[r].upload_raw(builds[r]) for r in range(world)]
        dp = [grp.ctxs[r].upload_raw(probes[r]) for r in range(world)]
        mm = [2 * npb] * world
        out = [grp.ctxs[r].alloc(2 * mm[r], np.uint32) for r in range(world)]
        hot = [grp.ctxs[r].alloc(2 * npb, np.uint32) for r in range(world)]
        if s >= 1.0:   # without the heavy-hitter path the hot keys' owner overflows its regions: reported, not mis-joined
            with pytest.raises(g.OlapError):
                grp.join(db, [nb] * world, dp, [npb] * world, out, mm)
        counts, hot_counts = grp.join(db, [nb] * world, dp, [npb] * world, out, mm, hot_pairs=hot, max_hot_pairs=[npb] * world)
        assert sum(counts) + sum(hot_counts) == exp_c
        if s >= 1.0:
            assert sum(hot_counts) > 0.3 * exp_c       # the heavy hitters carry a large share of the probe side
        boff = [r * nb for r in range(world)]; poff = [r * npb for r in range(world)]
        got = []
        for r in range(world):
            l, rt = grp.pairs_to_global(r, out[r], counts[r], boff, poff); got.append(np.stack([l, rt], 1))
            l, rt = grp.hot_pairs_to_global(r, hot[r], hot_counts[r], boff, poff); got.append(np.stack([l, rt], 1))
        got = np.concatenate(got).astype(np.uint32)
        assert (O.sort_pairs(got) == O.sort_pairs(exp_pairs)).all()


# ------------------------------------------------------------------------------------------------- sharded GROUP BY (C ABI)
def _sharded_group_by(world, key_shards, vi_shards, vf_shards, funcs, hint, out_capacity=None):
    import gpu_olap_engine_b200 as g
    from gpu_olap_engine_b200.shard import ShardGroup
    with ShardGroup([0] * world, 1024, 1024, 1) as grp:
        dk = [grp.ctxs[r].upload_raw(key_shards[r]) if len(key_shards[r]) else None for r in range(world)]
        di = [grp.ctxs[r].upload_raw(vi_shards[r]) if len(vi_shards[r]) else None for r in range(world)]
        df = [grp.ctxs[r].upload_raw(vf_shards[r]) if len(vf_shards[r]) else None for r in range(world)]
        aggs = [[(f, (df[r] if t == g.F64 else di[r]), t) for f, t in funcs] for r in range(world)]
        return grp.group_by(dk, [len(k) for k in key_shards], aggs, hint, out_capacity)


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world,G", [(1, 300), (2, 1000), (4, 50_000), (8, 7)])
def test_sharded_group_by_in_process(world, G):
    # olap_sharded_group_by_i64: local partials -> owner rank by key hash -> merge; == the oracle's group-by over all rows
    import gpu_olap_engine_b200 as g
    rng = np.random.default_rng(world * 1000 + G)
    sizes = [int(x) for x in rng.integers(50_000, 400_000, world)]
    if world >= 4:
        sizes[1] = 0                                                        # a rank without rows
    keys = [rng.integers(-G // 2, G - G // 2, n).astype(np.int64) for n in sizes]
    vi = [rng.integers(-(1 << 20), 1 << 20, n).astype(np.int64) for n in sizes]
    vf = [(rng.integers(0, 1 << 20, n) / 1024.0).astype(np.float64) for n in sizes]     # dyadic: sums are order-independent
    funcs = [(g.SUM, g.I64), (g.COUNT, g.I64), (g.MIN, g.I64), (g.MAX, g.F64), (g.AVG, g.I64), (g.AVG, g.F64), (g.SUM, g.F64), (g.COUNT, g.F64)]
    res = _sharded_group_by(world, keys, vi, vf, funcs, G)
    ak, ai, af = np.concatenate(keys), np.concatenate(vi), np.concatenate(vf)
    ek, er = O.group_by_i64(ak, [af if t == g.F64 else ai for _, t in funcs], [f for f, _ in funcs])
    lib = g.load()
    bits = world.bit_length() - 1
    got_k = np.concatenate([r[1] for r in res])
    assert len(got_k) == len(ek) and len(np.unique(got_k)) == len(got_k)                 # every group on exactly one rank
    for r, (ng, k, cols) in enumerate(res):
        assert all(lib.olap_shuffle_dest_of(int(x), bits) == r for x in k[:200])          # ... the rank that owns its key
    order = np.argsort(got_k)
    assert (got_k[order] == ek).all()
    for i, (f, t) in enumerate(funcs):
        got = np.concatenate([r[2][i] for r in res])[order]
        if f == g.AVG:
            assert np.allclose(got, er[i], rtol=1e-12, atol=0), i
        else:
            assert got.dtype == er[i].dtype and (got == er[i]).all(), i


def test_sharded_group_by_capacity_and_errors():
    import gpu_olap_engine_b200 as g
    rng = np.random.default_rng(5)
    keys = [rng.integers(0, 5000, 100_000).astype(np.int64) for _ in range(2)]
    vi = [rng.integers(0, 100, 100_000).astype(np.int64) for _ in range(2)]
    with pytest.raises(g.OlapError) as ei:                                   # 5000 groups over 2 ranks do not fit 100 slots each
        _sharded_group_by(2, keys, vi, vi, [(g.SUM, g.I64)], 5000, out_capacity=[100, 100])
    assert ei.value.status == 6
    with pytest.raises(g.OlapError) as ei:
        _sharded_group_by(1, keys[:1], vi[:1], vi[:1], [(g.SUM, g.I64)] * 16, 5000)
    assert ei.value.status == 1
