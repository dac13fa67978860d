"""Sample I/O in the reference's wire formats (host-side, the sample tensor is already on the host
after run()):

  save_csv      ≙ io::csv::save_csv      (io/csv.rs:47-69):     header chain, observation, dim_0..dim_{d-1}
  save_arrow    ≙ io::arrow::save_arrow  (io/arrow.rs:53-117):  Arrow IPC file, chain:u32, observation:u32, dim_i:f64
  save_parquet  ≙ io::parquet::save_parquet (io/parquet.rs:49-131): same schema as a Parquet file

`data` is the [chains, samples, dim] array every sampler returns.
"""
import csv

import numpy as np


def _table(data):
    import pyarrow as pa
    a = np.asarray(data)
    if a.ndim != 3:
        raise ValueError("expected [chains, samples, dim]")
    c, n, d = a.shape
    cols = {"chain": pa.array(np.repeat(np.arange(c, dtype=np.uint32), n)),
            "observation": pa.array(np.tile(np.arange(n, dtype=np.uint32), c))}
    flat = a.reshape(c * n, d).astype(np.float64, copy=False)
    for i in range(d):
        cols["dim_%d" % i] = pa.array(np.ascontiguousarray(flat[:, i]))
    return pa.table(cols)


def save_csv(data, filename):
    a = np.asarray(data)
    c, n, d = a.shape
    with open(filename, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["chain", "observation"] + ["dim_%d" % i for i in range(d)])
        for ci in range(c):
            for oi in range(n):
                w.writerow([ci, oi] + [repr(v.item()) for v in a[ci, oi]])


def save_arrow(data, filename):
    import pyarrow as pa
    t = _table(data)
    with pa.OSFile(filename, "wb") as sink, pa.ipc.new_file(sink, t.schema) as writer:
        writer.write_table(t)


def save_parquet(data, filename):
    import pyarrow.parquet as pq
    pq.write_table(_table(data), filename)
