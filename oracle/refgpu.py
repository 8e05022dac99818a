"""ctypes front-end of oracle/_ref/libref_kernels.so: the reference's OWN CUDA kernels
(gpu-kernel-compiler/kernels/join_kernels.cuh) run on the GPU.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The library is built from the read-only reference checkout by oracle/Makefile (target `ref`) and travels to the GPU
box as a prebuilt file; /root/reference itself is never read at test time."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libref_kernels.so")
_lib = None


def available() -> bool:
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_SO)
        vp, u32, i32 = C.c_void_p, C.c_uint32, C.c_int
        _lib.ref_hash_vectors.argtypes = [vp, u32, vp, vp, u32, vp]
        _lib.ref_radix_partition.argtypes = [vp, vp, u32, i32, vp, vp, vp]
        _lib.ref_hash_join.argtypes = [vp, u32, vp, u32, u32, vp, vp, u32, vp]
        _lib.ref_radix_hash_join.argtypes = [vp, u32, vp, u32, vp, vp, u32, vp]
        _lib.ref_hash_aggregate_u64.argtypes = [vp, vp, u32, u32, vp]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data


def hash_vectors(keys64, keys32):
    k64 = np.ascontiguousarray(keys64, np.int64); k32 = np.ascontiguousarray(keys32, np.int32)
    o64 = np.empty(len(k64), np.uint32); o32 = np.empty(len(k32), np.uint32)
    assert lib().ref_hash_vectors(_p(k64), len(k64), _p(o64), _p(k32), len(k32), _p(o32)) == 0
    return o64, o32


def radix_partition(keys, radix_shift=0):
    k = np.ascontiguousarray(keys, np.int64)
    begin = np.empty(257, np.uint32); ok = np.empty_like(k); oi = np.empty(len(k), np.uint32)
    assert lib().ref_radix_partition(_p(k), None, len(k), radix_shift, _p(begin), _p(ok), _p(oi)) == 0
    return begin, ok, oi


def hash_join(build, probe, max_matches, radix=False):
    b = np.ascontiguousarray(build, np.int64); p = np.ascontiguousarray(probe, np.int64)
    left = np.empty(max(max_matches, 1), np.uint32); right = np.empty(max(max_matches, 1), np.uint32); cnt = C.c_uint32(0)
    if radix:
        rc = lib().ref_radix_hash_join(_p(b), len(b), _p(p), len(p), _p(left), _p(right), max_matches, C.byref(cnt))
    else:
        rc = lib().ref_hash_join(_p(b), len(b), _p(p), len(p), 0, _p(left), _p(right), max_matches, C.byref(cnt))
    assert rc == 0
    k = min(cnt.value, max_matches)
    return cnt.value, np.stack([left[:k], right[:k]], 1)


def hash_aggregate_u64(group_keys, values, num_groups):
    gk = np.ascontiguousarray(group_keys, np.uint32); v = np.ascontiguousarray(values, np.uint64)
    out = np.zeros(num_groups, np.uint64)
    assert lib().ref_hash_aggregate_u64(_p(gk), _p(v), len(gk), num_groups, _p(out)) == 0
    return out
