"""gpu-olap-engine_b200: B200-native (sm_100a) execution core for gpu-olap-engine's data-parallel hot path.

The compute path is the AOT-compiled CUDA library ``libolap_b200.so`` behind the C ABI in ``include/olap_b200.h``;
this package is the host-side mirror of the reference's operator / kernel-wrapper interface above it.
"""
from ._lib import LIB_PATH, OlapError, load  # noqa: F401
from .core import (AVG, COUNT, EQ, F64, GE, GT, I64, JOIN_AUTO, JOIN_BROADCAST, JOIN_RADIX, JOIN_SORT_MERGE, LE, LT, MAX, MIN,  # noqa: F401
                   NE, SUM, Context, DeviceBuffer, JoinBuild, Term, typed_agg_results)
from . import kernels, plan  # noqa: F401,E402
from .executor import GpuExecutor, execute  # noqa: F401,E402
from .plan import (Avg, BinaryExpr, Column, Count, EngineConfig, GpuAggregate, GpuFilter, GpuHashJoin, GpuProjection, GpuTableScan,  # noqa: F401,E402
                   JoinType, Literal, Max, Min, Sum)
