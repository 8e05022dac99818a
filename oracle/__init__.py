"""ctypes front-end of the CPU oracle (oracle/oracle.c).  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module -- as the checker or as the timed CPU baseline, never on the product path.  The product
(``gpu_olap_engine_b200``) does not import it and fails loudly when its CUDA library is missing.

Every function cites the reference lines it restates in oracle.c's header.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

EQ, NE, LT, LE, GT, GE = range(6)          # CmpOp order, gpu-kernel-compiler/src/codegen.rs:10
SUM, COUNT, MIN, MAX, AVG = range(5)       # AggFunc order, gpu-kernel-compiler/src/aggregate_kernel.rs:17-25
I64, F64 = 0, 1


def build(force: bool = False) -> str:
    """Compile oracle.c with gcc (and oracle/_ref when the reference checkout is present)."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _sig()
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _sig():
    L = _lib
    u64, i64, i32, vp = C.c_uint64, C.c_int64, C.c_int, C.c_void_p
    L.orc_splitmix64.restype = u64; L.orc_splitmix64.argtypes = [u64, u64]
    L.orc_perm.restype = u64; L.orc_perm.argtypes = [u64, u64, u64]
    for f in (L.orc_gen_uniform_i64, L.orc_gen_perm_i64, L.orc_gen_dyadic_f64):
        f.restype = None; f.argtypes = [vp, u64, u64, u64, u64]
    L.orc_gen_unit_f64.restype = None; L.orc_gen_unit_f64.argtypes = [vp, u64, u64, u64]
    L.orc_filter_mask_i64.restype = None; L.orc_filter_mask_i64.argtypes = [vp, vp, i64, u64, i32, vp]
    L.orc_mask_to_indices.restype = u64; L.orc_mask_to_indices.argtypes = [vp, u64, vp]
    L.orc_take_8b.restype = None; L.orc_take_8b.argtypes = [vp, vp, u64, vp]
    L.orc_project_i64.restype = None; L.orc_project_i64.argtypes = [vp, vp, i64, u64, C.c_char, vp]
    L.orc_hash_int32.restype = C.c_uint32; L.orc_hash_int32.argtypes = [C.c_int32]
    L.orc_hash_int64.restype = C.c_uint32; L.orc_hash_int64.argtypes = [i64]
    L.orc_partition_id.restype = C.c_uint32; L.orc_partition_id.argtypes = [i64, i32]
    L.orc_radix_partition_i64.restype = None; L.orc_radix_partition_i64.argtypes = [vp, vp, C.c_uint32, i32, vp, vp, vp]
    L.orc_ref_build.restype = None; L.orc_ref_build.argtypes = [vp, vp, C.c_uint32, vp, vp, C.c_uint32]
    L.orc_ref_probe.restype = u64
    L.orc_ref_probe.argtypes = [vp, vp, C.c_uint32, vp, vp, C.c_uint32, vp, vp, vp, u64]
    L.orc_ref_hash_aggregate_u64.restype = None
    L.orc_ref_hash_aggregate_u64.argtypes = [vp, vp, C.c_uint32, vp, C.c_uint32]
    L.orc_hash_join_i64.restype = i32; L.orc_hash_join_i64.argtypes = [vp, u64, vp, u64, vp, vp, u64, vp]
    L.orc_hash_join_i64_mt.restype = i32; L.orc_hash_join_i64_mt.argtypes = [vp, u64, vp, u64, vp, vp, u64, vp, i32]
    L.orc_identity_value.restype = i64; L.orc_identity_value.argtypes = [i32]
    L.orc_agg_global_init.restype = None; L.orc_agg_global_init.argtypes = [vp]
    L.orc_agg_global_update.restype = None; L.orc_agg_global_update.argtypes = [vp, vp, i32, u64]
    L.orc_agg_global_final.restype = C.c_double; L.orc_agg_global_final.argtypes = [vp, i32]
    L.orc_group_by_i64.restype = u64
    L.orc_group_by_i64.argtypes = [vp, u64, u64, vp, vp, vp, i32, u64, vp, vp]
    L.orc_group_by_i64_mt.restype = u64
    L.orc_group_by_i64_mt.argtypes = [vp, u64, u64, vp, vp, vp, u64, vp, vp, i32]
    L.orc_fnv1a64.restype = u64; L.orc_fnv1a64.argtypes = [C.c_char_p, u64]
    L.orc_widen_column.restype = i32; L.orc_widen_column.argtypes = [vp, i32, vp, u64, vp]


# ---------------------------------------------------------------------------------------------- generators
def splitmix64(seed: int, i: int) -> int:
    return lib().orc_splitmix64(seed, i)


def perm(x: int, domain: int, seed: int) -> int:
    return lib().orc_perm(x, domain, seed)


def gen_uniform_i64(n, seed, domain, row0=0):
    out = np.empty(n, np.int64); lib().orc_gen_uniform_i64(_p(out), n, seed, row0, domain); return out


def gen_perm_i64(n, seed, domain, row0=0):
    out = np.empty(n, np.int64); lib().orc_gen_perm_i64(_p(out), n, seed, row0, domain); return out


def gen_dyadic_f64(n, seed, domain, row0=0):
    out = np.empty(n, np.float64); lib().orc_gen_dyadic_f64(_p(out), n, seed, row0, domain); return out


def gen_unit_f64(n, seed, row0=0):
    out = np.empty(n, np.float64); lib().orc_gen_unit_f64(_p(out), n, seed, row0); return out


# ---------------------------------------------------------------------------------------------- filter
def filter_mask_i64(lhs, op, rhs=None, literal=0):
    """executor.rs:443-476.  rhs is a column or None (then ``literal`` is broadcast)."""
    lhs = np.ascontiguousarray(lhs, np.int64)
    rhs = None if rhs is None else np.ascontiguousarray(rhs, np.int64)
    mask = np.empty(len(lhs), np.uint8)
    lib().orc_filter_mask_i64(_p(lhs), _p(rhs), int(literal), len(lhs), op, _p(mask))
    return mask


def mask_to_indices(mask):
    mask = np.ascontiguousarray(mask, np.uint8)
    idx = np.empty(len(mask), np.uint64)
    k = lib().orc_mask_to_indices(_p(mask), len(mask), _p(idx))
    return idx[:k].copy()


def take(col, idx):
    """executor.rs:496-507 for Int64 / Float64 columns."""
    col = np.ascontiguousarray(col); assert col.dtype.itemsize == 8
    idx = np.ascontiguousarray(idx, np.uint64)
    out = np.empty(len(idx), col.dtype)
    lib().orc_take_8b(_p(col), _p(idx), len(idx), _p(out))
    return out


def project_i64(l, op: str, r=None, literal=0):
    l = np.ascontiguousarray(l, np.int64)
    r = None if r is None else np.ascontiguousarray(r, np.int64)
    out = np.empty(len(l), np.int64)
    lib().orc_project_i64(_p(l), _p(r), int(literal), len(l), op.encode()[:1], _p(out))
    return out


# ---------------------------------------------------------------------------------------------- hash / partition
def hash_int32(k): return lib().orc_hash_int32(int(np.int32(k)))
def hash_int64(k): return lib().orc_hash_int64(int(np.int64(k)))
def partition_id(k, shift=0): return lib().orc_partition_id(int(np.int64(k)), shift)


def radix_partition_i64(keys, row_ids=None, radix_shift=0):
    keys = np.ascontiguousarray(keys, np.int64)
    ids = None if row_ids is None else np.ascontiguousarray(row_ids, np.uint32)
    begin = np.empty(257, np.uint32); ok = np.empty_like(keys); oi = np.empty(len(keys), np.uint32)
    lib().orc_radix_partition_i64(_p(keys), _p(ids), len(keys), radix_shift, _p(begin), _p(ok), _p(oi))
    return begin, ok, oi


def ref_build_probe(build_keys, probe_keys, capacity=None, max_matches=None, build_ids=None, probe_ids=None,
                    build_keys_by_row=None):
    """join_kernels.cuh:80-166 run sequentially.  Returns (count, left_ids, right_ids)."""
    bk = np.ascontiguousarray(build_keys, np.int64); pk = np.ascontiguousarray(probe_keys, np.int64)
    nb, npb = len(bk), len(pk)
    capacity = capacity or max(1, int(1.5 * nb))           # README.md:114 sizing rule
    entries = np.zeros(max(nb, 1) * 3, np.uint32)
    buckets = np.full(capacity, 0xFFFFFFFF, np.uint32)
    bi = None if build_ids is None else np.ascontiguousarray(build_ids, np.uint32)
    pi = None if probe_ids is None else np.ascontiguousarray(probe_ids, np.uint32)
    lib().orc_ref_build(_p(bk), _p(bi), nb, _p(entries), _p(buckets), capacity)
    by_row = bk if build_keys_by_row is None else np.ascontiguousarray(build_keys_by_row, np.int64)
    cap = max_matches if max_matches is not None else 4 * max(npb, 1) + 1024
    left = np.empty(cap, np.uint32); right = np.empty(cap, np.uint32)
    cnt = lib().orc_ref_probe(_p(pk), _p(pi), npb, _p(entries), _p(buckets), capacity, _p(by_row), _p(left), _p(right), cap)
    if max_matches is None and cnt > cap:
        return ref_build_probe(build_keys, probe_keys, capacity, int(cnt), build_ids, probe_ids, build_keys_by_row)
    k = min(cnt, cap)
    return int(cnt), left[:k].copy(), right[:k].copy()


def ref_hash_aggregate_u64(group_keys_u32, values_u64, num_groups):
    gk = np.ascontiguousarray(group_keys_u32, np.uint32); v = np.ascontiguousarray(values_u64, np.uint64)
    out = np.zeros(num_groups, np.uint64)
    lib().orc_ref_hash_aggregate_u64(_p(gk), _p(v), len(gk), _p(out), num_groups)
    return out


# ---------------------------------------------------------------------------------------------- join
def hash_join_i64(build, probe, max_matches=None, threads=1, want_pairs=True):
    """Inner equi-join; returns (count, pairs[k,2] uint32 = (left_row, right_row)).

    threads == 1 is ``ref1`` (executor.rs:238-253 + the omitted probe); threads > 1 is ``refN``."""
    b = np.ascontiguousarray(build, np.int64); p = np.ascontiguousarray(probe, np.int64)
    cnt = C.c_uint64(0)
    if not want_pairs:
        fn = lib().orc_hash_join_i64 if threads == 1 else lib().orc_hash_join_i64_mt
        args = [_p(b), len(b), _p(p), len(p), None, None, 0, C.byref(cnt)] + ([threads] if threads > 1 else [])
        assert fn(*args) == 0
        return int(cnt.value), np.empty((0, 2), np.uint32)
    cap = max_matches if max_matches is not None else max(len(p), 1024)
    while True:
        left = np.empty(cap, np.uint32); right = np.empty(cap, np.uint32)
        if threads == 1:
            rc = lib().orc_hash_join_i64(_p(b), len(b), _p(p), len(p), _p(left), _p(right), cap, C.byref(cnt))
        else:
            rc = lib().orc_hash_join_i64_mt(_p(b), len(b), _p(p), len(p), _p(left), _p(right), cap, C.byref(cnt), threads)
        assert rc == 0
        if max_matches is not None or cnt.value <= cap:
            break
        cap = int(cnt.value)
    k = min(int(cnt.value), cap)
    return int(cnt.value), np.stack([left[:k], right[:k]], axis=1)


def sort_pairs(pairs):
    """Canonical order for comparison: lexicographic by (right_id, left_id)."""
    pairs = np.asarray(pairs, np.uint32).reshape(-1, 2)
    order = np.lexsort((pairs[:, 0], pairs[:, 1]))
    return pairs[order]


# ---------------------------------------------------------------------------------------------- aggregate
class GlobalAgg:
    """executor.rs:308-346 for one aggregate expression, fed batch by batch."""

    def __init__(self):
        self._st = (C.c_uint64 * 4)()
        lib().orc_agg_global_init(self._st)

    def update(self, col):
        col = np.ascontiguousarray(col)
        t = F64 if col.dtype == np.float64 else I64
        if t == I64:
            col = col.astype(np.int64, copy=False)
        lib().orc_agg_global_update(self._st, _p(col), t, len(col))

    def final(self, func) -> float:
        return lib().orc_agg_global_final(self._st, func)


def group_by_i64(keys, values, funcs, compensated=False, threads=1):
    """Grouped aggregation (completion rule).

    keys: int64 array or None (global).  values: list of arrays (int64 / float64), one per aggregate.
    Returns (sorted_keys[G], results) where results[i] is int64 (COUNT, or SUM/MIN/MAX of int64) or float64."""
    n = len(values[0]) if values else (0 if keys is None else len(keys))
    k = None if keys is None else np.ascontiguousarray(keys, np.int64)
    cols = [np.ascontiguousarray(v, np.float64 if np.asarray(v).dtype == np.float64 else np.int64) for v in values]
    na = len(funcs)
    types = (C.c_int * max(na, 1))(*[F64 if c.dtype == np.float64 else I64 for c in cols])
    fn = (C.c_int * max(na, 1))(*funcs)
    ptrs = (C.c_void_p * max(na, 1))(*[c.ctypes.data for c in cols])
    cap = 1 << 10
    while True:
        ok = np.empty(cap, np.int64); oa = np.empty(max(na, 1) * cap, np.uint64)
        if threads == 1:
            g = lib().orc_group_by_i64(_p(k), n, na, ptrs, types, fn, int(compensated), cap, _p(ok), _p(oa))
        else:
            g = lib().orc_group_by_i64_mt(_p(k), n, na, ptrs, types, fn, cap, _p(ok), _p(oa), threads)
        if g != 0xFFFFFFFFFFFFFFFF:
            break
        cap *= 8
        assert cap <= 1 << 31, "oracle group_by failed"
    res = []
    for i, f in enumerate(funcs):
        raw = oa[i * cap:i * cap + g].copy()
        is_f = f == AVG or (f != COUNT and cols[i].dtype == np.float64)
        res.append(raw.view(np.float64) if is_f else raw.view(np.int64))
    return ok[:g].copy(), res


def reference_compat_view(results):
    """executor.rs:347: the reference emits every aggregate as Float64."""
    return [np.asarray(r, np.float64) for r in results]


# ---------------------------------------------------------------------------------------------- arrow layout
def fnv1a64(s: bytes) -> int:
    return lib().orc_fnv1a64(s, len(s))


_WIDEN = {np.dtype(t): i for i, t in enumerate(
    [np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint16, np.uint32, np.uint64, np.float32, np.float64])}


def widen_column(arr, validity=None):
    """record_batch_convert.rs:35-110: every numeric column becomes packed 8-byte values; nulls -> 0."""
    arr = np.ascontiguousarray(arr)
    code = _WIDEN[arr.dtype]
    out = np.empty(len(arr), np.uint64)
    v = None if validity is None else np.ascontiguousarray(validity, np.uint8)
    assert lib().orc_widen_column(_p(arr), code, _p(v), len(arr), _p(out)) == 0
    return out.view(np.float64) if code >= 8 else out.view(np.int64)
