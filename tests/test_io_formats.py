"""Sample I/O wire formats (io/csv.rs, io/arrow.rs, io/parquet.rs): column names, order, types, row order, and the
reference's own golden strings for the CSV writer (io/csv.rs:160-219)."""
import csv

import numpy as np
import pytest

from general_mcmc_b200 import io as gio


def test_csv_reference_golden_strings(tmp_path):
    p = tmp_path / "a.csv"
    gio.save_csv(np.zeros((0, 0, 0), np.float32), str(p))                    # test_save_csv_empty_data
    assert open(p).read().strip() == "chain,observation"
    gio.save_csv(np.array([[[42.0]]]), str(p))                               # test_save_csv_single_chain_single_obs
    assert open(p).read().strip() == "chain,observation,dim_0\n0,0,42"
    gio.save_csv(np.array([[[1, 2], [3, 4]], [[10, 20], [30, 40]]]), str(p))  # test_save_csv_multi_chain (integers)
    assert open(p).read() == "chain,observation,dim_0,dim_1\n0,0,1,2\n0,1,3,4\n1,0,10,20\n1,1,30,40\n"
    # test_save_csv_tensor_data: f32 values print with their shortest f32 digits
    gio.save_csv_tensor(np.array([[[1.0, 2.0], [3.0, 4.0]], [[1.1, 2.1], [3.1, 4.1]]]), str(p))
    rows = list(csv.reader(open(p)))
    assert rows[1:] == [["0", "0", "1", "2"], ["0", "1", "3", "4"], ["1", "0", "1.1", "2.1"], ["1", "1", "3.1", "4.1"]]
    assert "\r" not in open(p, newline="").read()


def test_rust_display_of_floats():
    cases = {1e-7: "0.0000001", 1e21: "1000000000000000000000", 0.1: "0.1", -2.5: "-2.5", 3.0: "3", float("inf"): "inf",
             float("-inf"): "-inf", 1.0 / 3.0: "0.3333333333333333", 5e-324: "0." + "0" * 323 + "5"}
    for v, want in cases.items():
        assert gio.rust_display(np.float64(v)) == want
    assert gio.rust_display(np.float64("nan")) == "NaN"
    assert gio.rust_display(np.float32(0.1)) == "0.1" and gio.rust_display(np.float32(16777216.0)) == "16777216"
    assert gio.rust_display(-0.0) == "-0" and gio.rust_display(7) == "7"


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
        assert not any(t.schema.field(i).nullable for i in range(len(t.schema)))      # Field::new(.., false)
        assert t.num_rows == 6
        assert t.column("chain").to_pylist() == [0, 0, 0, 1, 1, 1]
        assert t.column("observation").to_pylist() == [0, 1, 2, 0, 1, 2]
        assert np.allclose(t.column("dim_3").to_numpy(), data[:, :, 3].ravel())


def test_parquet_tensor_is_observation_major(tmp_path):
    """save_parquet_tensor (io/parquet.rs:154-222): tensor [observations, chains, dim]; columns observation, chain, dim_i;
    rows walk observations in the outer loop.  The reference's own test (parquet.rs tests) uses this 2 x 2 x 2 tensor."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    tensor = np.array([[[1.0, 2.0], [3.0, 4.0]], [[1.1, 2.1], [3.1, 4.1]]], np.float32)     # [obs, chain, dim]
    f = str(tmp_path / "t.parquet")
    gio.save_parquet_tensor(tensor, f)
    t = pq.read_table(f)
    assert t.column_names == ["observation", "chain", "dim_0", "dim_1"]
    assert t.column("observation").to_pylist() == [0, 0, 1, 1]
    assert t.column("chain").to_pylist() == [0, 1, 0, 1]
    assert np.allclose(t.column("dim_0").to_numpy(), [1.0, 3.0, 1.1, 3.1])
    assert np.allclose(t.column("dim_1").to_numpy(), [2.0, 4.0, 2.1, 4.1])
    assert t.schema.field("observation").type == pa.uint32() and not t.schema.field("dim_1").nullable


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_device_column_export_matches_host(tmp_path, dtype):
    """gmcmc_export_columns: the columns built on the GPU (tiled transpose + widening) equal the host walk of the
    reference writers, in both row orders, from host and from device-resident tensors; ragged tile edges included."""
    import pyarrow.parquet as pq
    import general_mcmc_b200 as gm
    ctx = gm.default_context()
    rng = np.random.default_rng(3)
    for c, n, d in [(1, 1, 1), (3, 5, 2), (37, 41, 33), (70, 19, 100)]:
        x = rng.standard_normal((c, n, d)).astype(dtype)
        for order in (gio.CHAIN_MAJOR, gio.OBS_MAJOR):
            want = gio.columns_from_host(x, order)
            got = gio.columns_from_device(ctx, x, c, n, d, dtype, order, on_device=False)
            for a, b in zip(got, want):
                assert np.array_equal(a, b)
    # a sampler's device-resident tensor straight to Parquet in the tensor order
    q0 = rng.standard_normal((64, 6)).astype(dtype)
    s = gm.HMC(gm.IsotropicGaussian(1.0, 6), q0, 0.3, 4, seed=5, ctx=ctx)
    ptr = s.run_device(10, 2)
    host = gm.HMC(gm.IsotropicGaussian(1.0, 6), q0, 0.3, 4, seed=5, ctx=ctx).run(10, 2)
    f = str(tmp_path / "dev.parquet")
    gio.save_parquet_tensor(None, f, device=(ctx, ptr, (64, 10, 6), dtype))
    t = pq.read_table(f)
    assert t.column("observation").to_pylist()[:65] == [0] * 64 + [1]
    assert np.array_equal(t.column("dim_5").to_numpy().reshape(10, 64), host[:, :, 5].T.astype(np.float64))
    chain, obs, dims = gio.columns_from_device(ctx, ptr, 64, 10, 6, dtype, gio.CHAIN_MAJOR, chain_base=1000)
    assert chain[0] == 1000 and chain[-1] == 1063 and np.array_equal(dims[2].reshape(64, 10), host[:, :, 2].astype(np.float64))
