"""Thin Python mirror of the C ABI: a Context (one per GPU), DeviceBuffer columns and the kernel-level calls.

This is plumbing above include/olap_b200.h -- every operation here is one C-ABI call into the sm_100a library.
Names follow the reference's intended device seam (FilterParams / JoinParams / AggregateParams, u64 device
pointers; gpu-kernel-compiler/src/{filter,join,aggregate}_kernel.rs).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import AggSpec, FilterTerm, OlapError, check

EQ, NE, LT, LE, GT, GE = range(6)
SUM, COUNT, MIN, MAX, AVG = range(5)
I64, F64 = 0, 1
JOIN_AUTO, JOIN_RADIX, JOIN_SORT_MERGE, JOIN_BROADCAST = -1, 0, 1, 2

_SRC_TYPES = {np.dtype(t): i for i, t in enumerate(
    [np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint16, np.uint32, np.uint64, np.float32, np.float64])}


class DeviceBuffer:
    """A device allocation from the context's stream-ordered pool (SlabAllocator::allocate/free analogue)."""

    def __init__(self, ctx: "Context", nbytes: int, dtype=np.uint8, count: Optional[int] = None):
        self.ctx = ctx
        self.nbytes = int(nbytes)
        self.dtype = np.dtype(dtype)
        self.count = self.nbytes // self.dtype.itemsize if count is None else count
        p = C.c_void_p()
        check(ctx.lib.olap_pool_alloc(ctx.handle, self.nbytes, None, C.byref(p)))
        self.ptr = p.value
        self._owns = True

    def free(self):
        if getattr(self, "_owns", False) and self.ptr is not None and self.ctx.handle is not None:
            check(self.ctx.lib.olap_pool_free(self.ctx.handle, self.ptr, None))
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def __len__(self):
        return self.count

    def to_numpy(self, count: Optional[int] = None) -> np.ndarray:
        n = self.count if count is None else int(count)
        out = np.empty(n, self.dtype)
        if n:
            check(self.ctx.lib.olap_memcpy_d2h(self.ctx.handle, out.ctypes.data, self.ptr, n * self.dtype.itemsize, None))
        check(self.ctx.lib.olap_stream_synchronize(self.ctx.handle, None))
        return out

    def view(self, dtype) -> "DeviceBuffer":
        v = object.__new__(DeviceBuffer)
        v.ctx, v.nbytes, v.dtype, v.ptr = self.ctx, self.nbytes, np.dtype(dtype), self.ptr
        v.count = self.nbytes // v.dtype.itemsize
        v._base, v._owns = self, False  # keeps the owner alive; a view never frees
        return v


def _ptr(x) -> Optional[int]:
    """Device address of a DeviceBuffer / torch tensor / int / None."""
    if x is None:
        return None
    if isinstance(x, DeviceBuffer):
        return x.ptr
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    raise TypeError(f"cannot take a device pointer from {type(x)}")


@dataclass
class Term:
    """One conjunct of a filter: lhs <op> (rhs column | literal).  FilterPredicate analogue (codegen.rs:23-28)."""
    lhs: object
    op: int
    literal: object = 0
    rhs: object = None
    type: int = I64

    def pack(self) -> FilterTerm:
        if self.type == F64:
            bits = int(np.float64(self.literal).view(np.int64))
        else:
            bits = int(np.int64(self.literal))
        return FilterTerm(_ptr(self.lhs), _ptr(self.rhs), bits, int(self.op), int(self.type))


def _terms(terms: Optional[Sequence[Term]]):
    if not terms:
        return None, 0
    arr = (FilterTerm * len(terms))(*[t.pack() for t in terms])
    return arr, len(terms)


class Context:
    """One per GPU.  Mirrors EngineConfig's meaningful fields (gpu-olap-core/src/lib.rs:20-31)."""

    def __init__(self, device: int = 0, max_gpu_memory: int = 0, num_streams: int = 8):
        self.lib = _lib.load()
        h = C.c_void_p()
        check(self.lib.olap_ctx_create(device, max_gpu_memory, num_streams, C.byref(h)))
        self.handle = h.value
        self.device = device

    def close(self):
        if getattr(self, "handle", None) is not None:
            self.lib.olap_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ------------------------------------------------------------------ memory
    def alloc(self, count: int, dtype=np.int64) -> DeviceBuffer:
        dt = np.dtype(dtype)
        return DeviceBuffer(self, max(int(count), 0) * dt.itemsize, dt, int(count))

    def upload(self, arr, validity_bitmap: Optional[np.ndarray] = None, bit_offset: int = 0, want_validity: bool = False):
        """record_batch_to_gpu_buffers for one column: copy + widen on device to the 8-byte layout."""
        arr = np.ascontiguousarray(arr)
        code = _SRC_TYPES[arr.dtype]
        n = len(arr) - bit_offset
        out = self.alloc(n, np.float64 if code >= 8 else np.int64)
        valid = self.alloc(n, np.uint8) if want_validity else None
        bm = None if validity_bitmap is None else np.ascontiguousarray(validity_bitmap, np.uint8)
        check(self.lib.olap_column_upload(self.handle, arr.ctypes.data, n, code, None if bm is None else bm.ctypes.data,
                                          bit_offset, out.ptr, _ptr(valid), None))
        check(self.lib.olap_stream_synchronize(self.handle, None))  # host array may be released by the caller
        return (out, valid) if want_validity else out

    def upload_raw(self, arr) -> DeviceBuffer:
        arr = np.ascontiguousarray(arr)
        out = DeviceBuffer(self, arr.nbytes, arr.dtype, len(arr))
        if arr.nbytes:
            check(self.lib.olap_memcpy_h2d(self.handle, out.ptr, arr.ctypes.data, arr.nbytes, None))
        check(self.lib.olap_stream_synchronize(self.handle, None))
        return out

    def synchronize(self):
        check(self.lib.olap_ctx_synchronize(self.handle))

    def pool_stats(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(self.lib.olap_pool_stats(self.handle, C.byref(a), C.byref(b), C.byref(c)))
        return {"used": a.value, "high_water": b.value, "reserved": c.value}

    def generate(self, n: int, kind: int, seed: int, domain: int = 1, row0: int = 0, out=None, stream=None):
        out = out if out is not None else self.alloc(n, np.float64 if kind >= 2 else np.int64)
        check(self.lib.olap_generate(self.handle, _ptr(out), n, kind, seed, row0, domain, stream))
        return out

    def generate_fk(self, n: int, perm_seed: int, domain: int, fk_seed: int, zipf_s: float = 0.0, row0: int = 0, out=None, stream=None):
        """Foreign keys into a build side made by generate(kind=1, seed=perm_seed, domain): uniform or Zipf(zipf_s) over its rows."""
        out = out if out is not None else self.alloc(n, np.int64)
        check(self.lib.olap_generate_fk(self.handle, _ptr(out), n, perm_seed, domain, fk_seed, row0, float(zipf_s), stream))
        return out

    # ------------------------------------------------------------------ filter / projection
    def filter(self, terms: Sequence[Term], n_rows: int, want_mask=True, want_sel=True, stream=None):
        """Returns (count, mask DeviceBuffer|None, sel DeviceBuffer|None)."""
        arr, nt = _terms(terms)
        mask = self.alloc(n_rows, np.uint8) if want_mask else None
        sel = self.alloc(n_rows, np.uint32) if want_sel else None
        cnt = C.c_uint64()
        check(self.lib.olap_filter(self.handle, arr, nt, n_rows, _ptr(mask), _ptr(sel), None, C.byref(cnt), stream))
        return cnt.value, mask, sel

    def filter_i64(self, col, n_rows: int, op: int, literal: int, want_mask=True, want_sel=True, stream=None):
        mask = self.alloc(n_rows, np.uint8) if want_mask else None
        sel = self.alloc(n_rows, np.uint32) if want_sel else None
        cnt = C.c_uint64()
        check(self.lib.olap_filter_i64(self.handle, _ptr(col), n_rows, op, literal, _ptr(mask), _ptr(sel), None, C.byref(cnt), stream))
        return cnt.value, mask, sel

    def gather(self, col, sel, n_sel: int, dtype=np.int64, stream=None) -> DeviceBuffer:
        out = self.alloc(n_sel, dtype)
        check(self.lib.olap_gather_8b(self.handle, _ptr(col), _ptr(sel), n_sel, out.ptr, stream))
        return out

    def filter_compact(self, terms: Sequence[Term], n_rows: int, cols: Sequence, dtypes: Sequence, stream=None):
        arr, nt = _terms(terms)
        outs = [self.alloc(n_rows, dt) for dt in dtypes]
        ins = (C.c_void_p * max(len(cols), 1))(*[_ptr(c) for c in cols])
        ops = (C.c_void_p * max(len(cols), 1))(*[o.ptr for o in outs])
        cnt = C.c_uint64()
        check(self.lib.olap_filter_compact(self.handle, arr, nt, n_rows, ins, len(cols), ops, None, C.byref(cnt), stream))
        return cnt.value, outs

    def project_i64(self, lhs, op: str, n_rows: int, rhs=None, literal: int = 0, stream=None) -> DeviceBuffer:
        out = self.alloc(n_rows, np.int64)
        check(self.lib.olap_project_i64(self.handle, _ptr(lhs), _ptr(rhs), int(np.int64(literal)), op.encode()[:1], n_rows, out.ptr, stream))
        return out

    def project_f64(self, lhs, op: str, n_rows: int, rhs=None, literal: float = 0.0, stream=None) -> DeviceBuffer:
        out = self.alloc(n_rows, np.float64)
        check(self.lib.olap_project_f64(self.handle, _ptr(lhs), _ptr(rhs), float(literal), op.encode()[:1], n_rows, out.ptr, stream))
        return out

    # ------------------------------------------------------------------ join
    def join_unmatched_rows(self, pairs, n_pairs: int, side: int, n_rows: int, stream=None) -> np.ndarray:
        """Row ids of one join side (0 = left / build, 1 = right / probe) that appear in no pair, ascending -- what Left / Right /
        Full joins append NULL-extended (logical_plan.rs:52-58)."""
        out = self.alloc(max(n_rows, 1), np.uint32)
        cnt = C.c_uint64()
        check(self.lib.olap_join_unmatched_rows(self.handle, _ptr(pairs), n_pairs, side, n_rows, out.ptr, None, C.byref(cnt), stream))
        return out.to_numpy(cnt.value)

    def hash_join_i64(self, build_keys, n_build: int, probe_keys, n_probe: int, max_matches: Optional[int] = None,
                      strategy: int = JOIN_AUTO, out_pairs=None, want_pairs: bool = True, stream=None):
        """JoinKernel::join analogue.  Returns (count, pairs DeviceBuffer of uint32[2*k])."""
        if max_matches is None:
            max_matches = max(n_probe, 1)
        pairs = out_pairs
        if pairs is None and want_pairs:
            pairs = self.alloc(2 * max_matches, np.uint32)
        cnt = C.c_uint64()
        st = self.lib.olap_hash_join_i64(self.handle, _ptr(build_keys), n_build, _ptr(probe_keys), n_probe, _ptr(pairs),
                                         max_matches if pairs is not None else 0, None, C.byref(cnt), strategy, stream)
        if st == _lib.ERR_OVERFLOW and not want_pairs:
            st = _lib.OK
        if st == _lib.ERR_OVERFLOW:
            err = OlapError(st, self.lib.olap_last_error().decode())
            err.count = cnt.value
            raise err
        check(st)
        return cnt.value, pairs

    def radix_partition_i64(self, keys, n_rows: int, bits: int, row_ids=None, stream=None):
        ok = self.alloc(n_rows, np.int64)
        oi = self.alloc(n_rows, np.uint32)
        begin = self.alloc((1 << bits) + 1, np.uint32)
        check(self.lib.olap_radix_partition_i64(self.handle, _ptr(keys), _ptr(row_ids), n_rows, bits, ok.ptr, oi.ptr, begin.ptr, stream))
        return ok, oi, begin

    # ------------------------------------------------------------------ aggregate
    def group_by_i64(self, keys, n_rows: int, aggs: Sequence[tuple], n_groups_hint: int, out_capacity: Optional[int] = None,
                     terms: Optional[Sequence[Term]] = None, stream=None, progs=None, key_prog=None):
        """AggregateKernel::aggregate analogue.  aggs: [(func, values, value_type)].

        Returns (n_groups, out_keys DeviceBuffer, out_aggs DeviceBuffer [n_aggs * out_capacity], out_capacity)."""
        na = len(aggs)
        specs = (AggSpec * max(na, 1))(*[AggSpec(_ptr(v), int(f), int(t)) for f, v, t in aggs])
        cap = out_capacity if out_capacity is not None else max(int(n_groups_hint), 1)
        ok = self.alloc(cap, np.int64)
        oa = self.alloc(max(na, 1) * cap, np.uint64)
        arr, nt = _terms(terms)
        g = C.c_uint64()
        if key_prog is not None:  # computed Int64 group key (olap_group_by_expr_i64): keys must be None
            kp = make_prog(key_prog, I64)
            built = [None if pr is None else make_prog(pr, aggs[i][2]) for i, pr in enumerate(progs or [None] * na)]
            ptrs = (C.c_void_p * max(na, 1))(*[None if b is None else C.addressof(b) for b in built])
            check(self.lib.olap_group_by_expr_i64(self.handle, C.addressof(kp), specs, ptrs, na, n_rows, n_groups_hint, arr, nt, ok.ptr, oa.ptr, cap,
                                                  None, C.byref(g), stream))
        elif progs is None:
            check(self.lib.olap_group_by_i64(self.handle, _ptr(keys), specs, na, n_rows, n_groups_hint, arr, nt, ok.ptr, oa.ptr, cap,
                                             None, C.byref(g), stream))
        else:  # progs[i]: None, or [("c", column) | ("l", literal) | "+" | "-" | "*" | "/", ...] in reverse Polish notation
            built = [None if pr is None else make_prog(pr, aggs[i][2]) for i, pr in enumerate(progs)]
            ptrs = (C.c_void_p * max(na, 1))(*[None if b is None else C.addressof(b) for b in built])
            check(self.lib.olap_group_by_prog_i64(self.handle, _ptr(keys), specs, ptrs, na, n_rows, n_groups_hint, arr, nt, ok.ptr, oa.ptr, cap,
                                                  None, C.byref(g), stream))
        return g.value, ok, oa, cap

    # ------------------------------------------------------------------ profiling
    def set_profiling(self, on: bool):
        check(self.lib.olap_set_profiling(self.handle, int(on)))

    def last_phase_times(self) -> dict:
        names = C.c_char_p(); ms = C.POINTER(C.c_float)(); n = C.c_int32()
        check(self.lib.olap_last_phase_times(self.handle, C.byref(names), C.byref(ms), C.byref(n)))
        if n.value == 0:
            return {}
        out = {}
        for nm, t in zip(names.value.decode().split(";"), [ms[i] for i in range(n.value)]):
            out[nm] = out.get(nm, 0.0) + t
        return out

    def launch_count(self) -> int:
        return self.lib.olap_launch_count(self.handle)


class JoinBuild:
    """Build side prepared once, probed batch by batch (olap_join_build_create / olap_join_probe)."""

    def __init__(self, ctx: Context, build_keys, n_build: int, strategy: int = JOIN_AUTO, stream=None):
        self.ctx = ctx
        h = C.c_void_p()
        check(ctx.lib.olap_join_build_create(ctx.handle, _ptr(build_keys), n_build, strategy, stream, C.byref(h)))
        self.handle = h.value

    def probe(self, probe_keys, n_probe: int, probe_row_base: int = 0, out_pairs=None, max_matches: int = 0, stream=None) -> int:
        """Returns this batch's exact match count; pairs (left_row, probe_row_base + local row) go to out_pairs."""
        cnt = C.c_uint64()
        st = self.ctx.lib.olap_join_probe(self.handle, _ptr(probe_keys), n_probe, probe_row_base, _ptr(out_pairs),
                                          max_matches if out_pairs is not None else 0, None, C.byref(cnt), stream)
        if st == _lib.ERR_OVERFLOW:
            err = OlapError(st, self.ctx.lib.olap_last_error().decode()); err.count = cnt.value
            raise err
        check(st)
        return cnt.value

    def close(self):
        if getattr(self, "handle", None):
            self.ctx.lib.olap_join_build_destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_prog(items, value_type) -> "_lib.ExprProg":
    """[("c", DeviceBuffer) | ("l", number) | "+" | "-" | "*" | "/"] -> olap_expr_prog (literals take the aggregate's value type)."""
    from ._lib import ExprProg, PROG_MAX_OPS
    if not 1 <= len(items) <= PROG_MAX_OPS:
        raise ValueError("an expression program has 1..%d entries" % PROG_MAX_OPS)
    g = ExprProg()
    g.n_ops = len(items)
    for k, it in enumerate(items):
        if isinstance(it, str):
            g.op[k] = ord(it)
        elif it[0] == "c":
            g.op[k] = ord("c"); g.column[k] = _ptr(it[1])
        else:
            g.op[k] = ord("l")
            g.literal_bits[k] = int(np.float64(it[1]).view(np.int64)) if value_type == F64 else int(it[1])
    return g


def typed_agg_results(out_aggs: np.ndarray, cap: int, n_groups: int, aggs: Sequence[tuple]):
    """Split the agg-major 8-byte slots (aggregate_kernel.rs:98) into typed arrays."""
    res = []
    for i, (f, _, t) in enumerate(aggs):
        raw = out_aggs[i * cap:i * cap + n_groups]
        is_f = f == AVG or (f != COUNT and t == F64)
        res.append(raw.view(np.float64) if is_f else raw.view(np.int64))
    return res
