"""CPU tests of the multi-GPU range sharding's HOST-side arithmetic (no GPU): olap_shard_route_i64 runs the same geometry
(olap_shard_plan) and the same exact-division-by-multiplication routing (route_part / route_dest, csrc/djoin.cuh) as the shuffle
kernels, and is checked here against plain integer division.  The reference has no multi-GPU path (README.md:414); this follows
SURVEY.md 8(e): shard by key radix / range, every row to the rank that owns its key."""
import ctypes as C

import numpy as np
import pytest


def route(world, gmin, gmax, keys):
    import gpu_olap_engine_b200 as g
    lib = g.load()
    keys = np.ascontiguousarray(keys, np.int64)
    n = len(keys)
    rank = np.empty(n, np.int32); sub = np.empty(n, np.uint32); off = np.empty(n, np.uint32); geo = (C.c_uint64 * 4)()
    st = lib.olap_shard_route_i64(world, gmin, gmax, keys.ctypes.data, n, rank.ctypes.data, sub.ctypes.data, off.ctypes.data, geo)
    return st, rank, sub, off, [int(x) for x in geo]


def expected(world, gmin, gmax, keys, geo):
    Q, F, subw, slice_ = geo
    o = keys.astype(object) - gmin                    # Python integers: no wrap-around in the reference arithmetic
    inside = np.array([0 <= x <= gmax - gmin for x in o])
    rank = np.array([int(x) // slice_ if ok else -1 for x, ok in zip(o, inside)], np.int64)
    sub = np.array([(int(x) % slice_) // subw if ok else 0 for x, ok in zip(o, inside)], np.int64)
    off = np.array([int(x) % slice_ if ok else 0 for x, ok in zip(o, inside)], np.int64)
    return rank, sub, off


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_geometry_tiles_the_domain(world):
    for grange in [1, 7, 1 << 15, (1 << 15) + 1, 100_000, 1 << 20, (1 << 27) + 12345, 1 << 29, (1 << 31) - 17]:
        if world == 1 and grange > (1 << 31):
            continue
        st, _, _, _, (Q, F, subw, slice_) = route(world, -5, grange - 6, np.zeros(0, np.int64))
        per_rank = -(-grange // world)
        pl = max(-(-per_rank // (1 << 15)), 1)                        # table ranges of 2^15 keys per rank
        if world * -(-pl // 256) > 256:                               # more than 256 regions per slab side: the plan refuses, the caller hash-shards
            assert st == 5, (world, grange)
            continue
        assert st == 0, (world, grange)
        assert 1 <= F <= 256 and Q >= 1 and world * Q <= 256          # regions per slab side, one partition counter per thread
        assert subw == F << 15 and slice_ == Q * subw                 # sub-ranges are whole multiples of the 2^15-key table ranges
        assert slice_ * world >= grange                               # the ranks' slices cover the domain ...
        assert slice_ * world < 1 << 32                               # ... in 32-bit offsets
        assert (slice_ - Q * (1 << 15)) * world < grange or slice_ == Q << 15   # and are not a whole table range wider than needed


@pytest.mark.parametrize("world,gmin,grange", [(2, 0, 1 << 29), (8, -12345, (1 << 31) - 3), (4, 1 << 40, 5_000_000), (1, -(1 << 62), 1 << 20),
                                               (3, 17, 999_999), (8, 0, 300)])
def test_route_matches_integer_division(world, gmin, grange):
    gmax = gmin + grange - 1
    rng = np.random.default_rng(world * 31 + grange % 1000)
    st, _, _, _, geo = route(world, gmin, gmax, np.zeros(0, np.int64))
    assert st == 0
    Q, F, subw, slice_ = geo
    edges = []
    for k in range(0, world * Q + 1):                                  # every sub-range boundary, both sides
        edges += [gmin + k * subw - 1, gmin + k * subw, gmin + k * subw + 1]
    keys = np.concatenate([np.array([e for e in edges if gmin - 3 <= e <= gmax + 3], np.int64),
                           np.array([gmin - 1, gmin, gmax, gmax + 1], np.int64),
                           gmin + rng.integers(0, grange, 200_000, dtype=np.int64)])
    st, rank, sub, off, _ = route(world, gmin, gmax, keys)
    assert st == 0
    erank, esub, eoff = expected(world, gmin, gmax, keys, geo)
    assert (rank == erank).all() and (sub == esub).all() and (off == eoff).all()
    inside = rank >= 0
    assert int(rank.max()) < world and (sub[inside] < Q).all() and (off[inside] < slice_).all()
    # the offset a row carries plus its owner's slice start gives the key back: nothing is lost in the packed 8-byte word
    assert ((off[inside].astype(np.int64) + rank[inside].astype(np.int64) * slice_ + gmin) == keys[inside]).all()


def test_exact_division_by_multiplication_over_every_reachable_width():
    # route_part computes (offset >> 15) / F as a multiply-shift, route_dest partition / Q likewise: sweep the table-range count of a
    # one-rank domain so that every F in 1..256 and every Q in 1..256 occurs, and probe both sides of every multiple of F
    seen_f, seen_q = set(), set()
    for pl in list(range(1, 600)) + [1000, 4095, 4096, 4097, 20_000, 32_767, 40_001, 65_535, 65_536]:
        grange = pl << 15
        st, _, _, _, (Q, F, subw, slice_) = route(1, 0, grange - 1, np.zeros(0, np.int64))
        assert st == 0
        seen_f.add(F); seen_q.add(Q)
        k = np.arange(0, Q + 1, dtype=np.int64) * subw
        keys = np.concatenate([k - 1, k, k + (1 << 15) - 1])
        keys = keys[(keys >= 0) & (keys < grange)]
        st, rank, sub, off, _ = route(1, 0, grange - 1, keys)
        assert st == 0 and (rank == 0).all()
        assert (sub.astype(np.int64) == keys // subw).all() and (off.astype(np.int64) == keys).all(), (pl, Q, F)
    assert seen_f >= set(range(1, 257)) and {1, 2, 3, 79, 157, 256} <= seen_q


def test_route_refuses_what_the_plan_refuses():
    import gpu_olap_engine_b200 as g
    lib = g.load()
    st, *_ = route(8, 0, (1 << 32) - 100, np.zeros(0, np.int64))     # 8 ranks x 64 sub-ranges do not fit 256 regions
    assert st == 5 and b"too wide" in lib.olap_last_error()
    st, *_ = route(2, 10, 5, np.zeros(0, np.int64))
    assert st == 1
    st, *_ = route(9, 0, 100, np.zeros(0, np.int64))
    assert st == 1


def hot_table(keys_plus1, probe_plus1):
    import gpu_olap_engine_b200 as g
    lib = g.load()
    k = np.ascontiguousarray(keys_plus1, np.uint32); p = np.ascontiguousarray(probe_plus1, np.uint32)
    tab = np.zeros(8192, np.uint32); slots = np.empty(len(p), np.uint32); placed = C.c_uint32()
    st = lib.olap_shard_hot_table_host(k.ctypes.data, len(k), tab.ctypes.data, C.byref(placed), p.ctypes.data, len(p), slots.ctypes.data)
    return st, tab, placed.value, slots


@pytest.mark.parametrize("n_keys,dense", [(0, False), (1, False), (500, False), (2048, False), (2048, True)])
def test_heavy_hitter_table_builder_and_lookup_agree(n_keys, dense):
    # the 2048-bucket x 4-slot table of the Zipf path (csrc/djoin.cu): every key that was placed is found at its slot, nothing else is found,
    # slots fill in order inside a bucket (the lookup's "second bucket only when the first is full" rule depends on it)
    rng = np.random.default_rng(n_keys + dense)
    keys = (np.arange(n_keys, dtype=np.uint32) * 7 + 1) if dense else (rng.choice(1 << 31, n_keys, replace=False).astype(np.uint32) + 1)
    others = np.setdiff1d(rng.integers(1, 1 << 32, 50_000, dtype=np.uint64).astype(np.uint32), keys)
    others = others[(others != 0) & (others != 0xFFFFFFFF)]
    st, tab, placed, slots = hot_table(np.concatenate([keys, keys[:10]]) if n_keys else keys, np.concatenate([keys, others]))   # duplicates are dropped
    if n_keys + min(n_keys, 10) > 2048:
        assert st == 1                                                  # more than a quarter of the slots: refused
        st, tab, placed, slots = hot_table(keys, np.concatenate([keys, others]))
    assert st == 0
    stored = tab[tab != 0]
    assert len(stored) == placed == len(np.unique(stored)) and set(stored.tolist()) <= set(keys.tolist())
    assert placed >= n_keys - 2                                         # both buckets full is rare at a quarter of the capacity
    found = slots[:n_keys]
    was_placed = np.isin(keys, stored)
    assert (found[was_placed] != 0xFFFFFFFF).all() and (tab[found[was_placed]] == keys[was_placed]).all()
    assert (found[~was_placed] == 0xFFFFFFFF).all() and (slots[n_keys:] == 0xFFFFFFFF).all()
    b = tab.reshape(2048, 4)
    assert ((b[:, 1:] == 0) | (b[:, :-1] != 0)).all()                   # no hole before a taken slot


def test_heavy_hitter_table_same_layout_for_any_key_order():
    # every rank must end up with the same slot for the same key whatever order the host language hands the keys over in
    rng = np.random.default_rng(3)
    keys = rng.choice(1 << 30, 2000, replace=False).astype(np.uint32) + 1
    _, t1, _, _ = hot_table(keys, keys[:1])
    _, t2, _, _ = hot_table(rng.permutation(keys), keys[:1])
    assert (t1 == t2).all()
