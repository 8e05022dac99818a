"""GPU parity of the dense-key group-by regime (csrc/groupby.cu run_dense_range): key ranges of 2^12 keys aggregated in
direct-address shared-memory tables behind a one- or two-level range partition.  Oracle: oracle/oracle.c orc_group_by_i64
(accumulator semantics executor.rs:308-346, identities aggregate_kernel.rs:27-33), groups compared sorted by key; i64 bit-exact,
f64 SUM/AVG bit-exact on dyadic data."""
import numpy as np
import pytest

import oracle as O
from test_gpu_parity import ALL5, check_group_by

pytestmark = pytest.mark.gpu


def phases_of(ctx):
    return ctx.last_phase_times()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("G,hint", [(10_000, 10_000),        # 3 key ranges: many CTAs per range, merged through the global direct table
                                    (10_000, 20_000),
                                    (300_000, 300_000),      # 74 ranges: one partition level, one CTA per range
                                    (3_000_000, 3_000_000),  # 733 ranges: two partition levels
                                    (40_000_000, 40_000_000)])
def test_dense_keys_take_the_range_regime(ctx, G, hint):
    ctx.set_profiling(True)
    n = (1 << 22) + 13
    k = O.gen_uniform_i64(n, 4, G) - 777          # negative minimum
    v = O.gen_uniform_i64(n, 5, 1 << 20) - (1 << 19)
    check_group_by(ctx, k, [v] * 5, ALL5, hint)
    ph = phases_of(ctx)
    assert "partition" in ph and ("partition_coarse" in ph) == (G > (1 << 20)), ph
    f = O.gen_dyadic_f64(n, 6, 1 << 20)
    check_group_by(ctx, k, [f] * 5, ALL5, hint)   # f64: dyadic values, sums are order-independent -> bit-exact


@pytest.mark.timeout(600)
def test_dense_regime_holes_count_only_and_single_group_ranges(ctx):
    n = (1 << 21) + 1
    k = O.gen_uniform_i64(n, 4, 30_000) * 5 + 11            # holes: 4 of 5 keys absent, range = 5 x groups
    v = O.gen_uniform_i64(n, 5, 1 << 30)
    check_group_by(ctx, k, [v, v, v], [O.COUNT, O.SUM, O.MAX], 30_000)
    check_group_by(ctx, k, [v], [O.COUNT], 30_000)           # COUNT only: no value column is carried
    k2 = np.where(np.arange(n) % 2 == 0, 5, 5 + 200_000).astype(np.int64)   # two groups, 49 empty ranges between them
    check_group_by(ctx, k2, [v, v], [O.SUM, O.MIN], 210_000)


@pytest.mark.timeout(600)
def test_dense_regime_falls_back_on_skew_and_notices_a_stale_key_range(ctx):
    import gpu_olap_engine_b200 as g
    ctx.set_profiling(True)
    n = (1 << 21) + 77
    k = O.gen_uniform_i64(n, 4, 200_000); v = O.gen_uniform_i64(n, 5, 1 << 20)
    k[: int(n * 0.6)] = 4242                                # 60 % of the rows in one key: its range's region overflows -> other regimes
    np.random.default_rng(3).shuffle(k)
    check_group_by(ctx, k, [v] * 4, [O.SUM, O.COUNT, O.MIN, O.MAX], 200_000)
    # remembered min / max of a key column at an address whose contents change
    k1 = O.gen_uniform_i64(n, 4, 100_000)
    dk, dv = ctx.upload_raw(k1), ctx.upload_raw(v)
    aggs = [(g.SUM, dv, g.I64), (g.COUNT, dv, g.I64)]

    def run(keys):
        ng, ok, oa, cap = ctx.group_by_i64(dk, n, aggs, 100_000, out_capacity=1 << 18)
        gk = ok.to_numpy(ng); res = g.typed_agg_results(oa.to_numpy(), cap, ng, aggs)
        ek, eres = O.group_by_i64(keys, [v, v], [O.SUM, O.COUNT])
        order = np.argsort(gk)
        assert ng == len(ek) and (gk[order] == ek).all() and all((a[order] == b).all() for a, b in zip(res, eres))
    run(k1)
    run(k1)
    k2 = k1 + 5_000_000                                      # same buffer, every key outside the remembered range
    ctx.lib.olap_memcpy_h2d(ctx.handle, dk.ptr, k2.ctypes.data, n * 8, None); ctx.synchronize()
    run(k2)
