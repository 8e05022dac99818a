"""The reference's kernel-wrapper seam (gpu-kernel-compiler/src/{join,aggregate,filter}_kernel.rs, codegen.rs) with the
same type and method names, over the C ABI: u64 device pointers in, results through out-pointers."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from enum import IntEnum
from typing import List

from ._lib import AggSpec, check, load
from .core import Context


class JoinStrategy(IntEnum):        # join_kernel.rs:13-18
    RadixHashJoin = 0
    SortMergeJoin = 1
    BroadcastHashJoin = 2


class AggFunc(IntEnum):             # aggregate_kernel.rs:17-25
    Sum = 0
    Count = 1
    Min = 2
    Max = 3
    Avg = 4

    def identity_value(self) -> int:  # aggregate_kernel.rs:27-33
        return load().olap_agg_identity(int(self))


class CmpOp(IntEnum):               # codegen.rs:10
    Eq = 0
    Ne = 1
    Lt = 2
    Le = 3
    Gt = 4
    Ge = 5


@dataclass
class JoinParams:                   # join_kernel.rs:20-27
    build_key_ptr: int
    probe_key_ptr: int
    n_build: int
    n_probe: int
    out_pairs_ptr: int
    out_count_ptr: int
    max_matches: int = 0            # capacity of out_pairs (the reference passes it to the probe kernel, join_kernels.cuh:127)


@dataclass
class AggregateSpec:                # aggregate_kernel.rs:37-44
    func: AggFunc
    value_col_index: int
    output_col_index: int
    value_type: int = 0             # OLAP_I64 / OLAP_F64


@dataclass
class AggregateParams:              # aggregate_kernel.rs:47-58
    group_key_ptr: int
    value_ptrs: List[int]
    n_rows: int
    n_groups: int
    out_ptr: int
    out_keys_ptr: int = 0
    out_capacity: int = 0


@dataclass
class FilterPredicate:              # codegen.rs:23-28
    col_index: int
    op: CmpOp
    literal: int


@dataclass
class FilterParams:                 # filter_kernel.rs:17-28
    column_ptr: int
    n_rows: int
    out_mask_ptr: int
    grid_dim: int = 0               # accepted for source compatibility; the AOT kernel sizes its own grid
    block_dim: int = 0


def compute_launch_params(n_rows: int):   # filter_kernel.rs:75-79 / aggregate_kernel.rs:127-131
    g, b = C.c_uint32(), C.c_uint32()
    load().olap_compute_launch_params(n_rows, C.byref(g), C.byref(b))
    return g.value, b.value


class JoinKernel:
    def __init__(self, ctx: Context):
        self.ctx = ctx

    @staticmethod
    def choose_strategy(n_build: int) -> JoinStrategy:   # join_kernel.rs:71-77
        return JoinStrategy(load().olap_choose_join_strategy(n_build))

    def join(self, strategy: JoinStrategy, params: JoinParams) -> int:
        """Returns the exact match count (also written to out_count_ptr); raises OlapError(OVERFLOW) when pairs were truncated."""
        cnt = C.c_uint64()
        check(self.ctx.lib.olap_hash_join_i64(self.ctx.handle, params.build_key_ptr, params.n_build, params.probe_key_ptr, params.n_probe,
                                              params.out_pairs_ptr or None, params.max_matches, params.out_count_ptr or None, C.byref(cnt),
                                              int(strategy), None))
        return cnt.value


class AggregateKernel:
    def __init__(self, ctx: Context):
        self.ctx = ctx

    def aggregate(self, specs: List[AggregateSpec], params: AggregateParams) -> int:
        """Output slot of spec i: out_ptr + output_col_index * out_capacity * 8 (agg-major, aggregate_kernel.rs:98)."""
        order = sorted(specs, key=lambda s: s.output_col_index)
        arr = (AggSpec * max(len(order), 1))(*[AggSpec(params.value_ptrs[s.value_col_index] if s.func != AggFunc.Count else None,
                                                       int(s.func), s.value_type) for s in order])
        cap = params.out_capacity or params.n_groups
        g = C.c_uint64()
        check(self.ctx.lib.olap_group_by_i64(self.ctx.handle, params.group_key_ptr or None, arr, len(order), params.n_rows, params.n_groups, None, 0,
                                             params.out_keys_ptr or None, params.out_ptr, cap, None, C.byref(g), None))
        return g.value


class FilterKernel:
    """The per-predicate PTX JIT (KernelCodegen::emit_filter_ptx, codegen.rs:36-47) is replaced by one AOT kernel that
    takes (op, literal) as arguments: ``launch`` takes the predicate where the reference takes the PTX string."""

    def __init__(self, ctx: Context):
        self.ctx = ctx

    def launch(self, predicate: FilterPredicate, params: FilterParams) -> int:
        cnt = C.c_uint64()
        check(self.ctx.lib.olap_filter_i64(self.ctx.handle, params.column_ptr, params.n_rows, int(predicate.op), predicate.literal,
                                           params.out_mask_ptr, None, None, C.byref(cnt), None))
        return cnt.value

    compute_launch_params = staticmethod(compute_launch_params)
