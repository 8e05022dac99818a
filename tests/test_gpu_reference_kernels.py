"""Pins the oracle AND the new core against the reference's own CUDA kernels (join_kernels.cuh) executed on the
B200 through oracle/_ref/libref_kernels.so -- the only part of the reference's hot path that can run at all."""
import numpy as np
import pytest

import oracle as O
from oracle import refgpu

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refgpu.available(), reason="oracle/_ref was not built (no reference checkout)")]

from test_oracle_golden import HASH32, HASH64  # noqa: E402


def test_reference_device_hashes_pin_the_vectors():
    k64 = [k for k, _ in HASH64]; k32 = [k for k, _ in HASH32]
    o64, o32 = refgpu.hash_vectors(k64, k32)
    assert o64.tolist() == [e for _, e in HASH64] and o32.tolist() == [e for _, e in HASH32]
    rnd = O.gen_uniform_i64(10000, 1, 2**63) - 2**62
    o64, _ = refgpu.hash_vectors(rnd, [0])
    assert o64.tolist() == [O.hash_int64(int(k)) for k in rnd]


def test_reference_partition_kernel_vs_oracle_multisets():
    keys = O.gen_uniform_i64(200000, 2, 2**40)
    rb, rk, ri = refgpu.radix_partition(keys)
    ob, ok, oi = O.radix_partition_i64(keys)
    assert (rb == ob).all()
    for p in range(256):                      # order inside a partition is whatever the atomics gave: compare multisets
        assert (np.sort(ri[rb[p]:rb[p + 1]]) == np.sort(oi[ob[p]:ob[p + 1]])).all()
    assert (keys[ri] == rk).all()


@pytest.mark.parametrize("radix", [False, True])
@pytest.mark.parametrize("nb,npb,dom", [(1000, 5000, 300), (100000, 1000000, 100000), (4096, 4096, 17), (50000, 20000, 2**62)])
def test_reference_join_kernels_vs_oracle_vs_core(ctx, radix, nb, npb, dom):
    b = O.gen_uniform_i64(nb, 11, dom); p = O.gen_uniform_i64(npb, 12, dom)
    exp_c, exp_pairs = O.hash_join_i64(b, p)
    rc, rpairs = refgpu.hash_join(b, p, max(exp_c, 1), radix=radix)
    assert rc == exp_c and (O.sort_pairs(rpairs) == exp_pairs).all()          # oracle == reference kernels
    db, dp = ctx.upload_raw(b), ctx.upload_raw(p)
    cnt, pairs = ctx.hash_join_i64(db, nb, dp, npb, max_matches=max(exp_c, 1), strategy=0 if radix else 2)
    assert cnt == rc and (O.sort_pairs(pairs.to_numpy(2 * cnt).reshape(-1, 2)) == O.sort_pairs(rpairs)).all()  # core == reference


def test_reference_join_truncates_silently_but_counts_exactly():
    b = np.arange(1000, dtype=np.int64); p = np.arange(1000, dtype=np.int64)
    rc, rpairs = refgpu.hash_join(b, p, 10)
    assert rc == 1000 and len(rpairs) == 10                                    # join_kernels.cuh:155-160


def test_reference_aggregate_kernel_vs_oracle_vs_core(ctx):
    import gpu_olap_engine_b200 as g
    n, G = 300000, 1000
    gk = O.gen_uniform_i64(n, 4, G).astype(np.uint32); v = O.gen_uniform_i64(n, 5, 1 << 20).astype(np.uint64)
    ref = refgpu.hash_aggregate_u64(gk, v, G)                                  # direct addressing, SUM only (:216-240)
    assert (ref == O.ref_hash_aggregate_u64(gk, v, G)).all()
    keys, res = O.group_by_i64(gk.astype(np.int64), [v.astype(np.int64)], [O.SUM])
    assert (ref[keys] == res[0].view(np.uint64)).all()
    dk, dv = ctx.upload_raw(gk.astype(np.int64)), ctx.upload_raw(v.astype(np.int64))
    ng, ok, oa, cap = ctx.group_by_i64(dk, n, [(g.SUM, dv, g.I64)], G)
    got_k = ok.to_numpy(ng); got = oa.to_numpy()[:ng]
    assert ng == len(keys) and (ref[got_k] == got).all()
