"""CPU tests: pin the oracle against every known-answer vector the reference holds for this path
 and against independent numpy/pandas restatements.  No GPU."""
import numpy as np
import pytest

import oracle as O


def test_reference_filter_vector():
    # gpu-olap-core/src/executor.rs:553-570: [5, 15, 25] > 10 -> [false, true, true]
    assert O.filter_mask_i64([5, 15, 25], O.GT, literal=10).tolist() == [0, 1, 1]
    f = _golden("reference_test_vectors.json")["filter"]
    assert O.filter_mask_i64(f["values"], O.GT, literal=f["literal"]).tolist() == f["mask"]


def test_filter_ops_and_unknown_op():
    a = np.array([-3, 0, 7, 7, 2**62], np.int64); b = np.array([0, 0, 7, 8, -2**62], np.int64)
    exp = {O.EQ: a == b, O.NE: a != b, O.LT: a < b, O.LE: a <= b, O.GT: a > b, O.GE: a >= b}
    for op, e in exp.items():
        assert (O.filter_mask_i64(a, op, rhs=b) == e).all()
    assert not O.filter_mask_i64(a, 17, rhs=b).any()  # executor.rs:464: unknown op -> false


def test_filter_gather_preserves_order():
    col = np.arange(100, dtype=np.int64) * 3
    mask = O.filter_mask_i64(col, O.GE, literal=150)
    idx = O.mask_to_indices(mask)
    assert (O.take(col, idx) == col[col >= 150]).all()
    f = np.linspace(0, 1, 100)
    assert (O.take(f, idx) == f[col >= 150]).all()


def test_identity_values():
    # gpu-kernel-compiler/src/aggregate_kernel.rs:139-143
    L = O.lib()
    assert L.orc_identity_value(O.SUM) == 0 and L.orc_identity_value(O.COUNT) == 0 and L.orc_identity_value(O.AVG) == 0
    assert L.orc_identity_value(O.MIN) == 2**63 - 1 and L.orc_identity_value(O.MAX) == -2**63


def _golden(name):
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name)) as f:
        return json.load(f)


# tests/golden/reference_test_vectors.json: the known-answer vectors of the reference (file:line cited inside)
_REF = _golden("reference_test_vectors.json")
HASH32 = [tuple(v) for v in _REF["hash_int32"]["vectors"]]
HASH64 = [tuple(v) for v in _REF["hash_int64"]["vectors"]]


def test_reference_hash_vectors():
    # derived from join_kernels.cuh:26-41 in SURVEY.md 8(c); the -m gpu tests run the reference's own device functions
    for k, e in HASH32:
        assert O.hash_int32(k) == e
    for k, e in HASH64:
        assert O.hash_int64(k) == e
    assert [[k, O.partition_id(k)] for k, _ in _REF["partition_ids"]["vectors"]] == _REF["partition_ids"]["vectors"]


def test_reference_partition_skew_on_dense_keys():
    begin, ok, oi = O.radix_partition_i64(np.arange(2**20, dtype=np.int64))
    sizes = np.diff(begin.astype(np.int64))
    assert sizes.min() == 3935 and sizes.max() == 4279 and sizes.sum() == 2**20
    for p in (0, 100, 255):
        seg = ok[begin[p]:begin[p + 1]]
        assert all(O.partition_id(int(k)) == p for k in seg[:16])
    assert (np.sort(oi) == np.arange(2**20)).all() and (ok == oi).all()


def test_widening_vectors():
    # arrow-interop/src/record_batch_convert.rs:201-217: Int32 1 widens to i64 1, 5 rows -> 40 bytes
    w = O.widen_column(np.array([1, 2, 3, 4, 5], np.int32))
    assert w.dtype == np.int64 and w.nbytes == 40 and w[0] == 1
    assert O.widen_column(np.array([2**64 - 1], np.uint64))[0] == -1          # u64 as i64 reinterprets
    assert O.widen_column(np.array([-5], np.int8))[0] == -5
    assert O.widen_column(np.array([1.5], np.float32))[0] == 1.5
    assert (O.widen_column(np.array([7, 8], np.int16), validity=[1, 0]) == [7, 0]).all()  # null -> 0 (:103)


def test_fnv1a_constants():
    # record_batch_convert.rs:123-130; standard FNV-1a test vectors
    assert O.fnv1a64(b"") == 0xcbf29ce484222325
    assert O.fnv1a64(b"a") == 0xaf63dc4c8601ec8c
    assert O.fnv1a64(b"foobar") == 0x85944171f73967e8


def test_projection_semantics():
    a = np.array([1, -2**63, 7, 2**63 - 1, 5], np.int64); b = np.array([0, -1, 2, 1, -3], np.int64)
    assert O.project_i64(a, '/', b).tolist() == [0, -2**63, 3, 2**63 - 1, -1]  # x/0 -> 0, wrapping MIN/-1, trunc toward 0
    assert O.project_i64(a, '+', b).tolist()[3] == -2**63                       # wraps
    assert O.project_i64(a, '*', literal=2).tolist()[1] == 0
    assert not O.project_i64(a, '%', b).any()                                     # unknown op -> 0 (executor.rs:435)


def _np_join(b, p):
    import pandas as pd
    l = pd.DataFrame({"k": b, "l": np.arange(len(b), dtype=np.uint32)})
    r = pd.DataFrame({"k": p, "r": np.arange(len(p), dtype=np.uint32)})
    m = l.merge(r, on="k")
    return O.sort_pairs(np.stack([m.l.values, m.r.values], 1)) if len(m) else np.empty((0, 2), np.uint32)


@pytest.mark.parametrize("nb,npb,dom", [(0, 10, 5), (10, 0, 5), (1, 1, 1), (1000, 5000, 300), (5000, 1000, 100000), (4096, 4096, 17)])
def test_join_restatements_agree(nb, npb, dom):
    b = O.gen_uniform_i64(nb, 11, dom); p = O.gen_uniform_i64(npb, 12, dom)
    c1, p1 = O.hash_join_i64(b, p)
    cN, pN = O.hash_join_i64(b, p, threads=4)
    cR, l, r = O.ref_build_probe(b, p)                        # join_kernels.cuh build/probe, serialised
    exp = _np_join(b, p)
    assert c1 == cN == cR == len(exp)
    assert (O.sort_pairs(p1) == exp).all() and (O.sort_pairs(pN) == exp).all()
    assert (O.sort_pairs(np.stack([l, r], 1)) == exp).all()
    assert (p1 == O.sort_pairs(p1)).all()                     # ref1 emits in (right, left) order already


def test_join_edge_keys_and_truncation():
    b = np.array([-1, 2**32 + 1, 0, 2**63 - 1, -2**63, 2**32, 1], np.int64)   # hash_int64 fold collisions
    p = np.array([0, -1, 1, 2**32, 2**32 + 1, -2**63, 2**63 - 1, 99], np.int64)
    c, pairs = O.hash_join_i64(b, p)
    assert c == 7 and (b[pairs[:, 0]] == p[pairs[:, 1]]).all()
    cR, l, r = O.ref_build_probe(b, p)
    assert cR == 7
    c2, pairs2 = O.hash_join_i64(b, p, max_matches=3)         # count exact, pairs truncated (join_kernels.cuh:155-160)
    assert c2 == 7 and len(pairs2) == 3


def test_global_aggregate_reference_semantics():
    # executor.rs:308-346
    v = O.gen_uniform_i64(100000, 5, 1 << 20)
    g = O.GlobalAgg(); g.update(v[:40000]); g.update(v[40000:])
    assert g.final(O.SUM) == float(v.sum()) and g.final(O.COUNT) == 100000.0
    assert g.final(O.MIN) == float(v.min()) and g.final(O.MAX) == float(v.max())
    assert g.final(O.AVG) == float(v.sum()) / 100000.0
    e = O.GlobalAgg()
    assert e.final(O.SUM) == 0.0 and e.final(O.COUNT) == 0.0 and e.final(O.AVG) == 0.0
    assert e.final(O.MIN) == np.finfo(np.float64).max and e.final(O.MAX) == -np.finfo(np.float64).max


def test_group_by_against_pandas_and_compat_view():
    import pandas as pd
    k = O.gen_uniform_i64(200000, 4, 1000); v = O.gen_uniform_i64(200000, 5, 1 << 20); f = O.gen_dyadic_f64(200000, 6, 1 << 20)
    keys, res = O.group_by_i64(k, [v, v, v, v, v, f, f, f, f], [O.SUM, O.COUNT, O.MIN, O.MAX, O.AVG, O.SUM, O.MIN, O.MAX, O.AVG])
    g = pd.DataFrame({"k": k, "v": v, "f": f}).groupby("k").agg(s=("v", "sum"), c=("v", "count"), mn=("v", "min"), mx=("v", "max"),
                                                               av=("v", "mean"), fs=("f", "sum"), fmn=("f", "min"), fmx=("f", "max"), fav=("f", "mean"))
    assert (keys == g.index.values).all()
    for got, col in zip(res, ["s", "c", "mn", "mx", "av", "fs", "fmn", "fmx", "fav"]):
        exp = g[col].values
        assert np.allclose(got, exp, rtol=1e-15, atol=0) if got.dtype == np.float64 else (got == exp).all(), col
    assert (res[5] == g.fs.values).all()  # dyadic: bit-exact regardless of order
    # reference-compat view (executor.rs:347 emits Float64): identical while magnitudes stay below 2^53
    compat = O.reference_compat_view(res)
    gl = O.GlobalAgg(); gl.update(v)
    assert compat[0].sum() == gl.final(O.SUM)
    kN, rN = O.group_by_i64(k, [v, v, v, v], [O.SUM, O.COUNT, O.MIN, O.MAX], threads=4)
    assert (kN == keys).all() and all((a == b).all() for a, b in zip(rN, res[:4]))


def test_group_by_no_key_and_empty():
    keys, res = O.group_by_i64(None, [np.array([], np.int64)] * 2, [O.SUM, O.COUNT])
    assert keys.tolist() == [0] and res[0].tolist() == [0] and res[1].tolist() == [0]
    keys, res = O.group_by_i64(np.array([], np.int64), [np.array([], np.int64)], [O.SUM])
    assert len(keys) == 0


def test_compensated_sum_is_the_yardstick():
    f = O.gen_unit_f64(1 << 20, 9)
    k = np.zeros(1 << 20, np.int64)
    _, seq = O.group_by_i64(k, [f], [O.SUM])
    _, comp = O.group_by_i64(k, [f], [O.SUM], compensated=True)
    import math
    exact = math.fsum(f.tolist())
    assert comp[0][0] == exact
    assert abs(seq[0][0] - exact) / exact < 1e-12


def test_generators_are_counter_based_and_bijective():
    a = O.gen_uniform_i64(1000, 3, 77)
    b = np.concatenate([O.gen_uniform_i64(400, 3, 77), O.gen_uniform_i64(600, 3, 77, row0=400)])
    assert (a == b).all() and a.min() >= 0 and a.max() < 77
    for dom in (1, 2, 3, 1000, 1 << 12, 100000):
        p = O.gen_perm_i64(dom, 2, dom)
        assert (np.sort(p) == np.arange(dom)).all()
    d = O.gen_dyadic_f64(1000, 6, 1 << 20)
    assert (d * 1024 == np.round(d * 1024)).all()


# ------------------------------------------------------------------------------------------------- fixtures from the reference itself
def _ref_fixtures():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_kernels.json")) as f:
        return json.load(f)


def test_oracle_matches_the_reference_kernels_own_outputs():
    """tests/golden/ref_kernels.json holds what the reference's CUDA kernels (join_kernels.cuh, built unmodified into
    oracle/_ref) produced on a B200 for seeded inputs (tests/golden/make_ref_fixtures.py).  The oracle must reproduce every
    one of them on the CPU: hashes, partition sizes and membership, join counts and pair sets, direct-address sums."""
    fx = _ref_fixtures()
    assert [O.hash_int64(k) for k in fx["hash_int64"]["keys"]] == fx["hash_int64"]["hashes"]
    assert [O.hash_int32(k) for k in fx["hash_int32"]["keys"]] == fx["hash_int32"]["hashes"]
    begin, _, _ = O.radix_partition_i64(np.arange(1 << 20, dtype=np.int64))
    assert begin.tolist() == fx["partition_dense_2p20"]["begin"]
    sizes = np.diff(np.asarray(fx["partition_dense_2p20"]["begin"]))
    assert (sizes.min(), sizes.max()) == (3935, 4279)            # +-4 % skew of the reference hash on dense keys
    pr = fx["partition_random"]
    keys = O.gen_uniform_i64(pr["n"], pr["seed"], pr["domain"])
    begin, _, ids = O.radix_partition_i64(keys)
    assert begin.tolist() == pr["begin"]
    part_of_row = np.empty(pr["n"], np.int32)
    for p in range(256):
        part_of_row[ids[begin[p]:begin[p + 1]]] = p
    assert part_of_row[:200].tolist() == pr["partition_of_first_200_rows"]
    for j in fx["joins"]:
        b = O.gen_uniform_i64(j["nb"], j["seed_build"], j["domain"]); p = O.gen_uniform_i64(j["np"], j["seed_probe"], j["domain"])
        c, pairs = O.hash_join_i64(b, p)
        assert c == j["count"]
        sp = O.sort_pairs(pairs)
        chk = [int(sp[:, 0].astype(np.uint64).sum()), int(sp[:, 1].astype(np.uint64).sum()),
               int((sp[:, 0].astype(np.uint64) * 1000003 + sp[:, 1].astype(np.uint64)).sum() % (1 << 61))]
        assert chk == j["pairs_checksum"] and sp[:20].tolist() == j["first_pairs"]
    assert fx["join_truncation"] == {"count": 1000, "pairs_written": 10}   # join_kernels.cuh:155-160: exact count, silent truncation
    ha = fx["hash_aggregate_sum"]
    gk = O.gen_uniform_i64(ha["n"], ha["seed_keys"], ha["groups"]).astype(np.uint32)
    v = O.gen_uniform_i64(ha["n"], ha["seed_values"], 1 << 20).astype(np.uint64)
    assert O.ref_hash_aggregate_u64(gk, v, ha["groups"]).tolist() == ha["sums"]

