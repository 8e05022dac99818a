"""executor::execute (gpu-olap-core/src/executor.rs:15-24) over the C ABI: builds the C plan tree, runs it on the GPU and
returns pyarrow RecordBatches through the Arrow C Data Interface.  Also the in-memory table registry that makes
GpuTableScan real (Catalog::get_table_data, catalog.rs:83-87)."""
from __future__ import annotations

import ctypes as C
from typing import List

import pyarrow as pa

from . import _lib, plan as P
from ._lib import check
from .core import Context


class _CArrow:
    """Heap storage for one ArrowArray (80 B) + ArrowSchema (72 B) pair."""

    def __init__(self):
        self.array = C.create_string_buffer(80)
        self.schema = C.create_string_buffer(72)

    @property
    def array_ptr(self):
        return C.addressof(self.array)

    @property
    def schema_ptr(self):
        return C.addressof(self.schema)


class GpuExecutor:
    """Mirror of the reference's (private) GpuExecutor: one per query context (executor.rs:26-35)."""

    def __init__(self, config: P.EngineConfig = None, device: int = 0, ctx: Context = None):
        self.config = config or P.EngineConfig()
        self.ctx = ctx or Context(device, self.config.max_gpu_memory if ctx is None and config is not None else 0, self.config.num_streams)
        self.lib = self.ctx.lib

    # ---- catalog
    def register_table(self, table_name: str, batches) -> None:
        """Make RecordBatches (or a pyarrow Table) scannable as ``table_name``; columns are uploaded and widened on device."""
        if isinstance(batches, pa.Table):
            batches = batches.combine_chunks().to_batches() or [pa.RecordBatch.from_pylist([], schema=batches.schema)]
        if isinstance(batches, pa.RecordBatch):
            batches = [batches]
        for b in batches:
            c = _CArrow()
            b._export_to_c(c.array_ptr, c.schema_ptr)
            try:
                check(self.lib.olap_register_batch(self.ctx.handle, table_name.encode(), c.array_ptr, c.schema_ptr))
            finally:
                pa.Schema._import_from_c(c.schema_ptr)  # releases the exported schema

    def load_table(self, table_name: str, path: str, batch_size: int = 1 << 20):
        """Catalog::load_table (catalog.rs:32-73) wired to the device: read the Parquet file (cached below 10 M rows, streamed row group
        by row group above) and register every batch, so that GpuTableScan(table_name) has device-resident columns to scan."""
        from .catalog import Catalog
        if not hasattr(self, "catalog"):
            self.catalog = Catalog()
        md = self.catalog.load_table(table_name, path, batch_size)
        self.drop_table(table_name)
        for b in self.catalog.iter_batches(table_name, batch_size):
            self.register_table(table_name, b)
        return md

    def register_device_table(self, table_name: str, columns, n_rows: int) -> None:
        """Append one batch of device-resident columns ({name: (DeviceBuffer, OLAP_I64 | OLAP_F64)}) to ``table_name`` -- copied device
        to device, no host round trip (olap_register_device_batch)."""
        names = list(columns)
        n = len(names)
        cn = (C.c_char_p * n)(*[x.encode() for x in names])
        ct = (C.c_int32 * n)(*[int(columns[x][1]) for x in names])
        cp = (C.c_void_p * n)(*[columns[x][0].ptr for x in names])
        check(self.lib.olap_register_device_batch(self.ctx.handle, table_name.encode(), n, cn, ct, cp, None, n_rows))

    def drop_table(self, table_name: str) -> None:
        check(self.lib.olap_drop_table(self.ctx.handle, table_name.encode()))

    # ---- plan lowering
    def _expr(self, e):
        if isinstance(e, P.Column):
            return self.lib.olap_expr_column(e.index, e.name.encode())
        if isinstance(e, P.Literal):
            return self.lib.olap_expr_literal(str(e.value).encode())
        if isinstance(e, P.BinaryExpr):
            return self.lib.olap_expr_binary(self._expr(e.left), e.op.encode(), self._expr(e.right))
        raise TypeError(f"not a PhysicalExpr: {e!r}")

    def _exprs(self, es):
        arr = (C.c_void_p * max(len(es), 1))(*[self._expr(e) for e in es])
        return arr, len(es)

    def _plan(self, p):
        L = self.lib
        if isinstance(p, P.GpuTableScan):
            schema_ptr, keep = None, None
            if p.schema is not None:
                keep = _CArrow()
                p.schema._export_to_c(keep.schema_ptr)
                schema_ptr = keep.schema_ptr
            proj = None if p.projection is None else (C.c_int64 * max(len(p.projection), 1))(*p.projection)
            node = L.olap_plan_table_scan(p.table_name.encode(), schema_ptr, proj, -1 if p.projection is None else len(p.projection))
            if keep is not None:
                pa.Schema._import_from_c(keep.schema_ptr)  # release
            return node
        if isinstance(p, P.GpuFilter):
            return L.olap_plan_filter(self._plan(p.input), self._expr(p.predicate))
        if isinstance(p, P.GpuProjection):
            arr, n = self._exprs(p.exprs)
            return L.olap_plan_projection(self._plan(p.input), arr, n)
        if isinstance(p, P.GpuHashJoin):
            la, nl = self._exprs(p.left_keys); ra, nr = self._exprs(p.right_keys)
            return L.olap_plan_hash_join(self._plan(p.left), self._plan(p.right), la, nl, ra, nr, int(p.join_type))
        if isinstance(p, P.GpuAggregate):
            ga, ng = self._exprs(p.group_by)
            ia, na = self._exprs([a.input for a in p.aggr_exprs])
            funcs = (C.c_int32 * max(na, 1))(*[a.func for a in p.aggr_exprs])
            return L.olap_plan_aggregate(self._plan(p.input), ga, ng, funcs, ia, na)
        raise TypeError(f"not a PhysicalPlan: {p!r}")

    # ---- execute
    def execute(self, plan) -> List[pa.RecordBatch]:
        node = self._plan(plan)
        if not node:
            raise _lib.OlapError(_lib.ERR_BAD_ARG, "invalid physical plan")
        res = C.c_void_p()
        try:
            check(self.lib.olap_execute(self.ctx.handle, node, C.byref(res)))
        finally:
            self.lib.olap_plan_free(node)
        out = []
        try:
            for i in range(self.lib.olap_result_num_batches(res)):
                c = _CArrow()
                check(self.lib.olap_result_export(res, i, c.array_ptr, c.schema_ptr))
                out.append(pa.RecordBatch._import_from_c(c.array_ptr, c.schema_ptr))
        finally:
            self.lib.olap_result_free(res)
        return out


def execute(plan, config: P.EngineConfig = None, executor: GpuExecutor = None) -> List[pa.RecordBatch]:
    """``pub async fn execute(plan, config) -> Result<Vec<RecordBatch>>`` (executor.rs:15-24)."""
    return (executor or GpuExecutor(config)).execute(plan)
