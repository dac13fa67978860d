"""Sample I/O wire formats (io/csv.rs, io/arrow.rs, io/parquet.rs): column names, order, types and row order."""
import csv

import numpy as np

from general_mcmc_b200 import io as gio


def test_csv_arrow_parquet_schema_and_roundtrip(tmp_path):
    import pyarrow as pa
    import pyarrow.parquet as pq
    data = np.arange(2 * 3 * 4, dtype=np.float32).reshape(2, 3, 4) * 0.5
    p = tmp_path / "s.csv"
    gio.save_csv(data, str(p))
    rows = list(csv.reader(open(p)))
    assert rows[0] == ["chain", "observation", "dim_0", "dim_1", "dim_2", "dim_3"]
    assert rows[1][:2] == ["0", "0"] and rows[4][:2] == ["1", "0"] and len(rows) == 7
    assert float(rows[6][5]) == data[1, 2, 3]
    for name, reader in (("s.arrow", lambda f: pa.ipc.open_file(f).read_all()), ("s.parquet", pq.read_table)):
        f = str(tmp_path / name)
        (gio.save_arrow if name.endswith("arrow") else gio.save_parquet)(data, f)
        t = reader(f)
        assert t.column_names == ["chain", "observation", "dim_0", "dim_1", "dim_2", "dim_3"]
        assert t.schema.field("chain").type == pa.uint32() and t.schema.field("dim_0").type == pa.float64()
        assert t.num_rows == 6
        assert t.column("chain").to_pylist() == [0, 0, 0, 1, 1, 1]
        assert t.column("observation").to_pylist() == [0, 1, 2, 0, 1, 2]
        assert np.allclose(t.column("dim_3").to_numpy(), data[:, :, 3].ravel())
