"""CPU test of the Catalog::load_table mirror (gpu-olap-core/src/catalog.rs:32-99): Parquet file -> schema / location / row count /
in-memory batches below the 10 M-row threshold, streamed above it."""
import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest


def test_load_table_metadata_cache_and_streaming(tmp_path, monkeypatch):
    from gpu_olap_engine_b200 import catalog as cat
    t = pa.table({"id": np.arange(250_000, dtype=np.int64), "amount": np.arange(250_000, dtype=np.float64) / 4, "q": np.arange(250_000, dtype=np.int32) % 50})
    path = tmp_path / "orders.parquet"
    pq.write_table(t, path, row_group_size=60_000)
    c = cat.Catalog()
    md = c.load_table("orders", str(path), batch_size=100_000)
    assert md.row_count == 250_000 == c.get_row_count("orders") and c.get_table_location("orders") == str(path)
    assert c.get_schema("orders").names == ["id", "amount", "q"]
    data = c.get_table_data("orders")                                   # below the threshold: cached (catalog.rs:50-57)
    assert data is not None and sum(b.num_rows for b in data) == 250_000
    assert pa.Table.from_batches(data).equals(t)
    monkeypatch.setattr(cat, "CACHE_ROW_LIMIT", 1000)                   # above it: no cache (:58-61), batches are streamed from the file
    md2 = c.load_table("big", str(path), batch_size=100_000)
    assert md2.data_cache is None and c.get_table_data("big") is None
    assert pa.Table.from_batches(list(c.iter_batches("big", 100_000))).equals(t)
    with pytest.raises(KeyError, match="Table not found"):
        c.get_schema("missing")
    with pytest.raises(Exception):
        c.load_table("nope", str(tmp_path / "nope.parquet"))
