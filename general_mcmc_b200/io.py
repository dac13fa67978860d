"""Sample I/O in the reference's wire formats.

  save_csv / save_csv_tensor   ≙ io::csv::save_csv / save_csv_tensor       (io/csv.rs:47-69, 110-147)
  save_arrow                   ≙ io::arrow::save_arrow                     (io/arrow.rs:53-117)   Arrow IPC file
  save_parquet                 ≙ io::parquet::save_parquet                 (io/parquet.rs:49-131)
  save_parquet_tensor          ≙ io::parquet::save_parquet_tensor          (io/parquet.rs:154-222): the tensor is
                               [observation, chain, dim] and the columns come in the order observation, chain, dim_i

Columns: chain:u32, observation:u32, dim_0 .. dim_{d-1}:f64, all non-nullable, one row per (chain, observation) pair.
CSV values are printed the way Rust's `Display` prints them (shortest round-trip digits, never an exponent, no trailing
".0": 42.0 -> "42", 1e-7 -> "0.0000001"), records end with "\\n" (the `csv` crate's default terminator).

`data` is the [chains, samples, dim] array every sampler returns.  With `device=(ctx, ptr, dtype)` — the pointer
gmcmc_run_device returned — the columns are built on the GPU by gmcmc_export_columns (tiled transpose + widening to f64),
so a multi-GB sample tensor is never transposed on the host; `columns_from_device` exposes that step on its own.
"""
import ctypes as C

import numpy as np

from . import _lib as L

CHAIN_MAJOR, OBS_MAJOR = 0, 1


def rust_display(v):
    """`format!("{}", v)` of a Rust integer / f32 / f64 (the reference writes every CSV field with `to_string()`)."""
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    if np.isnan(v):
        return "NaN"
    if np.isinf(v):
        return "inf" if v > 0 else "-inf"
    if v == 0:
        return "-0" if np.signbit(v) else "0"
    return np.format_float_positional(v, unique=True, trim="-")


def columns_from_host(data, row_order=CHAIN_MAJOR):
    """chain, observation (uint32 [rows]) and dims (float64 [d, rows]) of a host [chains, samples, dim] array."""
    a = np.asarray(data)
    if a.ndim != 3:
        raise ValueError("expected [chains, samples, dim]")
    c, n, d = a.shape
    if row_order == CHAIN_MAJOR:
        chain = np.repeat(np.arange(c, dtype=np.uint32), n)
        obs = np.tile(np.arange(n, dtype=np.uint32), c)
        flat = a.reshape(c * n, d)
    else:
        chain = np.tile(np.arange(c, dtype=np.uint32), n)
        obs = np.repeat(np.arange(n, dtype=np.uint32), c)
        flat = a.transpose(1, 0, 2).reshape(c * n, d)
    return chain, obs, np.ascontiguousarray(flat.T.astype(np.float64))


def columns_from_device(ctx, dev_ptr, n_chains, n_samples, dim, dtype, row_order=CHAIN_MAJOR, chain_base=0, on_device=True):
    """The same columns built on the GPU from a device-resident [chains, samples, dim] tensor (gmcmc_export_columns)."""
    rows = int(n_chains) * int(n_samples)
    chain = np.empty(rows, np.uint32)
    obs = np.empty(rows, np.uint32)
    dims = np.empty((int(dim), rows), np.float64)
    src = C.c_void_p(dev_ptr) if on_device else L.ptr(dev_ptr)
    L.check(L.lib().gmcmc_export_columns(ctx._h, src, C.c_size_t(n_chains), C.c_size_t(n_samples), C.c_size_t(dim),
                                         L.dtype_code(dtype), 1 if on_device else 0, int(row_order), C.c_uint32(chain_base),
                                         L.ptr(chain), L.ptr(obs), L.ptr(dims)))
    return chain, obs, dims


def _columns(data, device, row_order):
    if device is not None:
        ctx, ptr, shape, dtype = device
        return columns_from_device(ctx, ptr, shape[0], shape[1], shape[2], dtype, row_order)
    return columns_from_host(data, row_order)


def _table(chain, obs, dims, first="chain"):
    import pyarrow as pa
    cols = [("chain", pa.array(chain, pa.uint32())), ("observation", pa.array(obs, pa.uint32()))]
    if first == "observation":
        cols.reverse()
    for i in range(dims.shape[0]):
        cols.append(("dim_%d" % i, pa.array(dims[i], pa.float64())))
    schema = pa.schema([pa.field(name, arr.type, nullable=False) for name, arr in cols])
    return pa.Table.from_arrays([arr for _, arr in cols], schema=schema)


def save_csv(data, filename, device=None):
    """≙ save_csv (io/csv.rs:47-69).  Integer arrays print as integers, floats as Rust's Display prints them."""
    a = None if device is not None else np.asarray(data)
    if a is not None and a.ndim == 3 and not np.issubdtype(a.dtype, np.floating):
        c, n, d = a.shape
        rows_iter = ((ci, oi, [str(int(v)) for v in a[ci, oi]]) for ci in range(c) for oi in range(n))
        n_dims = d
    else:
        chain, obs, dims = _columns(data, device, CHAIN_MAJOR)
        src_dtype = device[3] if device is not None else a.dtype
        vals = dims.astype(src_dtype, copy=False) if np.dtype(src_dtype) == np.float32 else dims
        n_dims = dims.shape[0]
        rows_iter = ((int(chain[r]), int(obs[r]), [rust_display(v) for v in vals[:, r]]) for r in range(chain.size))
    with open(filename, "w", newline="") as f:
        f.write(",".join(["chain", "observation"] + ["dim_%d" % i for i in range(n_dims)]) + "\n")
        for ci, oi, vals_s in rows_iter:
            f.write(",".join([str(ci), str(oi)] + vals_s) + "\n")


def save_csv_tensor(tensor, filename, device=None):
    """≙ save_csv_tensor (io/csv.rs:110-147): a [chains, observations, dim] tensor, values as f32."""
    if device is not None:
        return save_csv(None, filename, device=device)
    return save_csv(np.asarray(tensor, np.float32), filename)


def save_arrow(data, filename, device=None):
    import pyarrow as pa
    t = _table(*_columns(data, device, CHAIN_MAJOR))
    with pa.OSFile(filename, "wb") as sink, pa.ipc.new_file(sink, t.schema) as writer:
        writer.write_table(t)


def save_parquet(data, filename, device=None):
    import pyarrow.parquet as pq
    pq.write_table(_table(*_columns(data, device, CHAIN_MAJOR)), filename)


def save_parquet_tensor(tensor, filename, device=None):
    """≙ save_parquet_tensor (io/parquet.rs:154-222): `tensor` is [observations, chains, dim]; rows are observation-major
    and the columns are observation, chain, dim_0 ...  With `device=(ctx, ptr, (chains, samples, dim), dtype)` the
    sampler's own [chains, samples, dim] tensor is exported in that row order on the GPU (no host permute)."""
    import pyarrow.parquet as pq
    if device is not None:
        chain, obs, dims = _columns(None, device, OBS_MAJOR)
    else:
        a = np.asarray(tensor)
        if a.ndim != 3:
            raise ValueError("expected [observations, chains, dim]")
        chain, obs, dims = columns_from_host(a.transpose(1, 0, 2), OBS_MAJOR)
    pq.write_table(_table(chain, obs, dims, first="observation"), filename)
