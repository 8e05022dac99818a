"""Catalog::load_table mirror (gpu-olap-core/src/catalog.rs:32-73): a Parquet file becomes a named table -- schema, location,
row count, and, for tables below the reference's in-memory threshold (10 M rows, :50), the RecordBatches themselves.  Host-side
Python over pyarrow only; GpuExecutor.load_table hands the batches to olap_register_batch, which is where the GpuTableScan source of
SURVEY 8(f) rank 1 gets its device-resident columns.  Tables above the threshold are streamed row group by row group instead of being
cached on the host (the reference leaves them unread: data_cache = None, :58-61)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Iterator, List, Optional

import pyarrow as pa
import pyarrow.parquet as pq

CACHE_ROW_LIMIT = 10_000_000          # catalog.rs:50


@dataclass
class TableMetadata:                  # catalog.rs:15-22
    schema: pa.Schema
    location: str
    row_count: int
    data_cache: Optional[List[pa.RecordBatch]]


class Catalog:
    def __init__(self):
        self.tables: Dict[str, TableMetadata] = {}

    def load_table(self, table_name: str, path: str, batch_size: int = 1 << 20) -> TableMetadata:
        f = pq.ParquetFile(path)                                        # raises like File::open / try_new do (:35-39)
        rows = f.metadata.num_rows
        cache = list(f.iter_batches(batch_size=batch_size)) if rows < CACHE_ROW_LIMIT else None
        md = TableMetadata(f.schema_arrow, str(path), rows, cache)
        self.tables[table_name] = md
        return md

    def _get(self, table_name: str) -> TableMetadata:
        if table_name not in self.tables:
            raise KeyError(f"Table not found: {table_name}")           # catalog.rs:78, :85, :92, :99
        return self.tables[table_name]

    def get_schema(self, table_name: str) -> pa.Schema:
        return self._get(table_name).schema

    def get_table_data(self, table_name: str) -> Optional[List[pa.RecordBatch]]:
        return self._get(table_name).data_cache

    def get_table_location(self, table_name: str) -> str:
        return self._get(table_name).location

    def get_row_count(self, table_name: str) -> int:
        return self._get(table_name).row_count

    def iter_batches(self, table_name: str, batch_size: int = 1 << 20) -> Iterator[pa.RecordBatch]:
        """The table's RecordBatches: the cache when there is one, else streamed from the Parquet file row group by row group."""
        md = self._get(table_name)
        if md.data_cache is not None:
            yield from md.data_cache
        else:
            yield from pq.ParquetFile(md.location).iter_batches(batch_size=batch_size)
