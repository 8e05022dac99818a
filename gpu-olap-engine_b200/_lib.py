"""ctypes binding of include/olap_b200.h.  There is no fallback: if libolap_b200.so is missing the import fails."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libolap_b200.so")

OK, ERR_BAD_ARG, ERR_OOM, ERR_OVERFLOW, ERR_CUDA, ERR_UNSUPPORTED, ERR_CAPACITY = range(7)
_STATUS_NAMES = ["OK", "BAD_ARG", "OOM", "OVERFLOW", "CUDA", "UNSUPPORTED", "CAPACITY"]


class OlapError(RuntimeError):
    """Mirror of the anyhow::Error the Rust wrapper would raise; carries the C status code."""

    def __init__(self, status: int, message: str):
        super().__init__(f"olap_b200 status {status} ({_STATUS_NAMES[status] if 0 <= status < 7 else '?'}): {message}")
        self.status = status
        self.message = message


class FilterTerm(C.Structure):
    _fields_ = [("lhs", C.c_void_p), ("rhs", C.c_void_p), ("literal_bits", C.c_int64), ("op", C.c_int32), ("type", C.c_int32)]


class AggSpec(C.Structure):
    _fields_ = [("values", C.c_void_p), ("func", C.c_int32), ("value_type", C.c_int32)]


PROG_MAX_OPS = 10


class ExprProg(C.Structure):
    """olap_expr_prog: reverse-Polish arithmetic over device columns / literals, evaluated inside the aggregation kernel."""
    _fields_ = [("n_ops", C.c_int32), ("op", C.c_int8 * PROG_MAX_OPS), ("column", C.c_void_p * PROG_MAX_OPS),
                ("literal_bits", C.c_int64 * PROG_MAX_OPS)]


# name -> (restype, argtypes); the test-suite checks that every OLAP_API symbol of the header is listed here
vp, u64, i64, i32, u32 = C.c_void_p, C.c_uint64, C.c_int64, C.c_int32, C.c_uint32
pu64, pvp = C.POINTER(C.c_uint64), C.POINTER(C.c_void_p)
SIGNATURES = {
    "olap_ctx_create": (i32, [i32, u64, i32, pvp]),
    "olap_ctx_destroy": (i32, [vp]),
    "olap_ctx_synchronize": (i32, [vp]),
    "olap_ctx_stream": (vp, [vp, i32]),
    "olap_last_error": (C.c_char_p, []),
    "olap_version": (C.c_char_p, []),
    "olap_pool_alloc": (i32, [vp, u64, vp, pvp]),
    "olap_pool_free": (i32, [vp, vp, vp]),
    "olap_pool_stats": (i32, [vp, pu64, pu64, pu64]),
    "olap_host_alloc": (i32, [u64, pvp]),
    "olap_host_free": (i32, [vp]),
    "olap_memcpy_h2d": (i32, [vp, vp, vp, u64, vp]),
    "olap_memcpy_d2h": (i32, [vp, vp, vp, u64, vp]),
    "olap_stream_synchronize": (i32, [vp, vp]),
    "olap_column_upload": (i32, [vp, vp, u64, i32, vp, u64, vp, vp, vp]),
    "olap_filter": (i32, [vp, C.POINTER(FilterTerm), i32, u64, vp, vp, vp, pu64, vp]),
    "olap_filter_i64": (i32, [vp, vp, u64, i32, i64, vp, vp, vp, pu64, vp]),
    "olap_gather_8b": (i32, [vp, vp, vp, u64, vp, vp]),
    "olap_filter_compact": (i32, [vp, C.POINTER(FilterTerm), i32, u64, pvp, i32, pvp, vp, pu64, vp]),
    "olap_project_i64": (i32, [vp, vp, vp, i64, C.c_char, u64, vp, vp]),
    "olap_project_f64": (i32, [vp, vp, vp, C.c_double, C.c_char, u64, vp, vp]),
    "olap_hash_join_i64": (i32, [vp, vp, u64, vp, u64, vp, u64, vp, pu64, i32, vp]),
    "olap_join_build_create": (i32, [vp, vp, u64, i32, vp, pvp]),
    "olap_join_probe": (i32, [vp, vp, u64, u32, vp, u64, vp, pu64, vp]),
    "olap_join_build_destroy": (i32, [vp]),
    "olap_join_gather_8b": (i32, [vp, vp, vp, i32, u64, vp, vp]),
    "olap_join_unmatched_rows": (i32, [vp, vp, u64, i32, u64, vp, vp, pu64, vp]),
    "olap_choose_join_strategy": (i32, [u64]),
    "olap_radix_partition_i64": (i32, [vp, vp, vp, u64, i32, vp, vp, vp, vp]),
    "olap_partition_of": (u32, [i64, i32]),
    "olap_shuffle_partition_i64": (i32, [vp, vp, vp, u64, i32, vp, vp, vp, vp, vp]),
    "olap_shuffle_dest_of": (u32, [i64, i32]),
    "olap_shuffle_scratch_words": (u64, [vp, i32]),
    "olap_shuffle_count_i64": (i32, [vp, vp, u64, i32, vp, vp, vp]),
    "olap_shuffle_scatter_i64": (i32, [vp, vp, u64, i32, vp, vp, vp, vp, vp]),
    "olap_shard_create": (i32, [vp, i32, i32, u64, u64, i32, pvp]),
    "olap_shard_destroy": (i32, [vp]),
    "olap_shard_slab_bytes": (u64, [vp]),
    "olap_shard_attach": (i32, [vp, pvp]),
    "olap_shard_barrier": (i32, [vp, vp]),
    "olap_shard_signal": (i32, [vp, vp]),
    "olap_shard_wait": (i32, [vp, vp]),
    "olap_shard_key_range": (i32, [vp, vp, u64, C.POINTER(i64), C.POINTER(i64), vp]),
    "olap_shard_plan": (i32, [vp, i64, i64, u64, i32, vp]),
    "olap_shard_scatter": (i32, [vp, i32, i32, vp, u64, u32, vp]),
    "olap_shard_send": (i32, [vp, i32, i32, vp]),
    "olap_shard_build": (i32, [vp, vp]),
    "olap_shard_probe": (i32, [vp, i32, vp, u64, vp]),
    "olap_shard_finish": (i32, [vp, pu64, pu64, C.POINTER(u32), vp]),
    "olap_shard_hot_candidates": (i32, [vp, vp, u64, C.POINTER(u32), C.POINTER(u32), u32, C.POINTER(u32), vp]),
    "olap_shard_hot_set": (i32, [vp, C.POINTER(u32), u32, vp]),
    "olap_shard_hot_resolve": (i32, [vp, vp, vp]),
    "olap_shard_hot_finalize": (i32, [vp, vp, vp, u64, vp]),
    "olap_shard_pairs_to_global": (i32, [vp, vp, u64, vp, vp, vp, vp, vp]),
    "olap_shard_layout": (i32, [vp, pu64]),
    "olap_shard_sent_rows": (i32, [vp, pu64, vp]),
    "olap_shard_hot_table_host": (i32, [vp, u32, vp, vp, vp, u32, vp]),
    "olap_shard_route_i64": (i32, [i32, i64, i64, vp, u64, vp, vp, vp, vp]),
    "olap_shard_ctx_create_all": (i32, [i32, C.POINTER(i32), u64, u64, u64, i32, pvp, pvp]),
    "olap_sharded_hash_join_i64": (i32, [pvp, i32, pvp, pu64, pvp, pu64, pvp, pu64, pu64, pvp, pu64, pu64]),
    "olap_shard_hot_table": (i32, [vp, pu64, vp]),
    "olap_group_by_i64": (i32, [vp, vp, C.POINTER(AggSpec), i32, u64, u64, C.POINTER(FilterTerm), i32, vp, vp, u64, vp, pu64, vp]),
    "olap_group_by_prog_i64": (i32, [vp, vp, C.POINTER(AggSpec), vp, i32, u64, u64, C.POINTER(FilterTerm), i32, vp, vp, u64, vp, pu64, vp]),
    "olap_group_by_expr_i64": (i32, [vp, vp, C.POINTER(AggSpec), vp, i32, u64, u64, C.POINTER(FilterTerm), i32, vp, vp, u64, vp, pu64, vp]),
    "olap_sharded_group_by_i64": (i32, [pvp, i32, pvp, pvp, i32, pu64, u64, pvp, pvp, pu64, pu64]),
    "olap_agg_identity": (i64, [i32]),
    "olap_compute_launch_params": (None, [u64, C.POINTER(u32), C.POINTER(u32)]),
    "olap_expr_column": (vp, [u64, C.c_char_p]),
    "olap_expr_literal": (vp, [C.c_char_p]),
    "olap_expr_binary": (vp, [vp, C.c_char_p, vp]),
    "olap_expr_free": (None, [vp]),
    "olap_plan_table_scan": (vp, [C.c_char_p, vp, C.POINTER(C.c_int64), i32]),
    "olap_plan_filter": (vp, [vp, vp]),
    "olap_plan_projection": (vp, [vp, pvp, i32]),
    "olap_plan_hash_join": (vp, [vp, vp, pvp, i32, pvp, i32, i32]),
    "olap_plan_aggregate": (vp, [vp, pvp, i32, C.POINTER(i32), pvp, i32]),
    "olap_plan_free": (None, [vp]),
    "olap_register_batch": (i32, [vp, C.c_char_p, vp, vp]),
    "olap_register_device_batch": (i32, [vp, C.c_char_p, i32, C.POINTER(C.c_char_p), C.POINTER(i32), pvp, pvp, u64]),
    "olap_drop_table": (i32, [vp, C.c_char_p]),
    "olap_execute": (i32, [vp, vp, pvp]),
    "olap_result_num_batches": (i32, [vp]),
    "olap_result_num_rows": (i64, [vp, i32]),
    "olap_result_export": (i32, [vp, i32, vp, vp]),
    "olap_result_free": (None, [vp]),
    "olap_generate": (i32, [vp, vp, u64, i32, u64, u64, u64, vp]),
    "olap_generate_fk": (i32, [vp, vp, u64, u64, u64, u64, u64, C.c_double, vp]),
    "olap_set_profiling": (i32, [vp, i32]),
    "olap_last_phase_times": (i32, [vp, C.POINTER(C.c_char_p), C.POINTER(C.POINTER(C.c_float)), C.POINTER(i32)]),
    "olap_launch_count": (u64, [vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the sm_100a CUDA library has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C gpu-olap-engine_b200/csrc`). "
                "There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status: int) -> None:
    if status != OK:
        raise OlapError(status, load().olap_last_error().decode(errors="replace"))
