"""PhysicalPlan / PhysicalExpr / AggregateExpr / JoinType / EngineConfig -- the reference's operator API types
(gpu-olap-core/src/physical_plan.rs:11-84, logical_plan.rs:52-58, lib.rs:19-43) as Python dataclasses with the same
names and field meanings, so plans are built exactly as a reference user builds them (the operator API
has to be driven by constructing PhysicalPlan values directly)."""
from __future__ import annotations

from dataclasses import dataclass, field
from enum import IntEnum
from typing import List, Optional, Union


class JoinType(IntEnum):          # logical_plan.rs:52-58
    Inner = 0
    Left = 1
    Right = 2
    Full = 3


@dataclass
class EngineConfig:               # lib.rs:19-43 (same defaults)
    max_gpu_memory: int = 8 * 1024 * 1024 * 1024
    num_streams: int = 8
    use_unified_memory: bool = True     # accepted and ignored: stream-ordered device pools replace managed memory
    batch_size: int = 1024 * 1024
    enable_cache: bool = True


# ---- PhysicalExpr (physical_plan.rs:66-75)
@dataclass
class Column:
    index: int
    name: str


@dataclass
class Literal:
    value: str


@dataclass
class BinaryExpr:
    left: "PhysicalExpr"
    op: str
    right: "PhysicalExpr"


PhysicalExpr = Union[Column, Literal, BinaryExpr]


# ---- AggregateExpr (physical_plan.rs:77-84); func codes follow AggFunc (aggregate_kernel.rs:17-25)
@dataclass
class Sum:
    input: PhysicalExpr
    func = 0


@dataclass
class Count:
    input: PhysicalExpr
    func = 1


@dataclass
class Min:
    input: PhysicalExpr
    func = 2


@dataclass
class Max:
    input: PhysicalExpr
    func = 3


@dataclass
class Avg:
    input: PhysicalExpr
    func = 4


AggregateExpr = Union[Sum, Count, Min, Max, Avg]


# ---- PhysicalPlan (physical_plan.rs:11-64)
@dataclass
class GpuTableScan:
    table_name: str
    schema: object = None                       # pyarrow.Schema (Arc<Schema>); needed only when the table has no data registered
    projection: Optional[List[int]] = None


@dataclass
class GpuFilter:
    input: "PhysicalPlan"
    predicate: PhysicalExpr


@dataclass
class GpuProjection:
    input: "PhysicalPlan"
    exprs: List[PhysicalExpr]


@dataclass
class GpuHashJoin:
    left: "PhysicalPlan"
    right: "PhysicalPlan"
    left_keys: List[PhysicalExpr]
    right_keys: List[PhysicalExpr]
    join_type: JoinType = JoinType.Inner
    schema: object = None                       # derived by the executor; the reference's placeholder is Schema::empty()


@dataclass
class GpuAggregate:
    input: "PhysicalPlan"
    group_by: List[PhysicalExpr] = field(default_factory=list)
    aggr_exprs: List[AggregateExpr] = field(default_factory=list)
    schema: object = None


PhysicalPlan = Union[GpuTableScan, GpuFilter, GpuProjection, GpuHashJoin, GpuAggregate]
