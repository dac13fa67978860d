"""Edge cases through the C ABI: single chain, single dimension, zero collected samples, ragged grids, state
continuity (positions / set_positions / step), error reporting."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import general_mcmc_b200 as gm  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    return gm.default_context()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_single_chain_single_dim_and_empty_runs(ctx, oracle, dtype):
    tgt = gm.IsotropicGaussian(1.0, 1)
    q0 = np.array([[0.3]], dtype)
    mom = np.array([[[0.7]], [[-1.1]], [[0.2]]], dtype)
    ln_u = np.log(np.array([[0.5], [0.9], [0.1]], dtype))
    ref = oracle.hmc_run(tgt.kind, tgt.params(), q0, 0.2, 3, mom, ln_u)
    s = gm.HMC(tgt, q0, 0.2, 3, seed=1, ctx=ctx).set_math_mode(True)
    s.inject(mom, ln_u)
    out = s.run(3, 0)
    assert np.array_equal(out, ref["samples"])
    assert s.run(0, 0).shape == (1, 0, 1)            # nothing to do
    before = s.positions()
    s.run(0, 5)                                       # burn-in only: state moves, nothing returned
    assert s.positions().shape == (1, 1) and np.isfinite(s.positions()).all()
    s.set_positions(before)
    assert np.array_equal(s.positions(), before)
    s.step()
    assert s.counters().transitions == 3 + 5 + 1

    m = gm.MetropolisHastings(tgt, gm.IsotropicGaussian(0.5), np.array([[0.0]], dtype), ctx=ctx).seed(3)
    o = m.run(5, 2)
    assert o.shape == (1, 5, 1) and o.dtype == np.float64 and np.isfinite(o).all()


def test_ragged_grids_match_oracle(ctx, oracle):
    """Chain counts that are not multiples of the warp / CTA / staging sizes, dims that need padding lanes."""
    # the last three shapes have unpadded rows (d a multiple of the 16-byte unit): the flat state load / store of K1 with a
    # ragged last warp, in both dtypes
    for Cn, d, dtype in [(1, 7, np.float32), (33, 5, np.float32), (129, 37, np.float32), (1000, 3, np.float32),
                         (13, 8, np.float32), (77, 100, np.float32), (5, 16, np.float32), (77, 100, np.float64), (13, 6, np.float64)]:
        rng = np.random.default_rng(Cn * 100 + d)
        q0 = (1.0 + 0.2 * rng.standard_normal((Cn, d))).astype(dtype)
        mom = rng.standard_normal((2, Cn, d)).astype(dtype)
        ln_u = np.log(rng.random((2, Cn))).astype(dtype)
        ref = oracle.hmc_run(oracle.ROSENBROCK_ND, [], q0, 0.01, 4, mom, ln_u)
        s = gm.HMC(gm.RosenbrockND(d), q0, 0.01, 4, seed=1, ctx=ctx).set_math_mode(True)
        s.inject(mom, ln_u)
        assert np.array_equal(s.run(2, 0), ref["samples"]), (Cn, d, dtype)
        assert np.array_equal(s.positions(), ref["samples"][:, -1]), (Cn, d, dtype)      # the state written back at the launch end
        # fast mode (the production instantiation) takes the same load / store path
        f = gm.HMC(gm.RosenbrockND(d), q0, 0.01, 4, seed=1, ctx=ctx)
        f.inject(mom[:1], ln_u[:1])
        out = f.run(1, 0)
        tol = 1e-5 if dtype == np.float32 else 1e-10
        assert np.allclose(out[:, 0], ref["samples"][:, 0], rtol=tol, atol=tol), (Cn, d, dtype)
        assert np.array_equal(f.positions(), out[:, 0]), (Cn, d, dtype)


def test_stats_minimal_and_unsupported_are_reported(ctx):
    x = np.random.default_rng(0).standard_normal((2, 4, 1)).astype(np.float32)
    rhat, ess = gm.split_rhat_mean_ess(x, ctx)
    assert rhat.shape == (1,) and np.isfinite(rhat).all()
    with pytest.raises(gm.GmcmcError) as e:
        gm.split_rhat_mean_ess(np.zeros((2, 3, 1), np.float32), ctx)     # n < 4
    assert e.value.status == 1
    with pytest.raises(gm.GmcmcError):
        gm.HMC(gm.RosenbrockND(5), np.zeros((4, 5), np.float32), -0.1, 3, ctx=ctx)   # bad step size
    with pytest.raises(gm.GmcmcError):
        gm.MetropolisHastings(gm.GaussianMixture([1.0], np.zeros((1, 3))), gm.IsotropicGaussian(1.0), np.zeros((4, 3)), ctx=ctx)


def test_pipelined_host_run_equals_device_run(ctx):
    """gmcmc_run cuts big runs into chain chunks that overlap compute and the device->host copy; the samples are
    the ones the one-kernel device run produces."""
    Cn, d = 20000, 100
    q0 = (1.0 + 0.1 * np.random.default_rng(5).standard_normal((Cn, d))).astype(np.float32)
    a = gm.HMC(gm.RosenbrockND(d), q0, 0.005, 4, seed=42, ctx=ctx).run(3, 1)          # pipelined (>= 16384 chains)
    b = np.concatenate([gm.HMC(gm.RosenbrockND(d), q0[i:i + 5000], 0.005, 4, seed=42, ctx=ctx, chain_offset=i).run(3, 1)
                        for i in range(0, Cn, 5000)])
    assert np.array_equal(a, b)


# -------------------------------------------------------------------------------------------------
# integer-state Metropolis-Hastings (tests/metrohast_poisson_test.rs)
# -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,mk,kind,params,d,x0max", [
    ("poisson", lambda: gm.PoissonTarget(4.0), 0, [4.0], 1, 9),
    ("binomial", lambda: gm.BinomialTarget(10, 0.3), 1, [10, 0.3], 1, 10),
    ("poisson3d", lambda: gm.PoissonTarget(2.5), 0, [2.5], 3, 6),
])
def test_mh_int_bit_exact_vs_oracle(ctx, oracle, name, mk, kind, params, d, x0max):
    """MetropolisHastings<S = i32, T = f64> with injected proposal directions and ln u: decisions, log ratios, traces and
    final states bit-identical to the oracle's restatement of metropolis_hastings.rs:306-318 on the reference's discrete
    targets (ln k! summed term by term, clamped +-1 walk, ln 0.5 proposal terms kept in the ratio)."""
    Cn, n, n_discard = 257, 60, 9
    rng = np.random.default_rng(7)
    x0 = rng.integers(0, x0max + 1, size=(Cn, d)).astype(np.int32)
    steps = (rng.integers(0, 2, size=(n, Cn, d)) * 2 - 1).astype(np.int8)
    ln_u = np.log(rng.random((n, Cn)))
    ref = oracle.mh_int_run(kind, params, x0, steps, ln_u)
    s = gm.MetropolisHastings(mk(), gm.RandomWalkProposal(), x0, ctx=ctx)
    s.inject_int(steps, ln_u)
    out = s.run(n - n_discard, n_discard)
    diag = s.diagnostics()
    assert out.dtype == np.float64
    assert np.array_equal(diag["accepted"], ref["accepted"])
    assert np.array_equal(diag["log_accept"], ref["log_ratio"])
    assert np.array_equal(out, ref["samples"][:, n_discard:, :])
    assert np.array_equal(s.positions(), ref["x"])
    assert 0.2 < ref["accepted"].mean() < 0.98


def test_mh_int_poisson_and_binomial_distributions(ctx):
    """The reference's own acceptance test (tests/metrohast_poisson_test.rs:92-140, 254-290): 20,000 draws after 2,000
    burn-in, empirical frequencies of k = 0..10 within 0.05 of the pmf — here over 4,096 chains, so within 0.005."""
    from math import comb, exp, factorial
    Cn = 4096
    s = gm.MetropolisHastings(gm.PoissonTarget(4.0), gm.RandomWalkProposal(), np.zeros((Cn, 1), np.int32), ctx=ctx).seed(42)
    out = s.run(2000, 2000)
    assert np.array_equal(out, np.round(out)) and out.min() >= 0
    freq = np.bincount(out.astype(np.int64).ravel(), minlength=11)[:11] / out.size
    pmf = np.array([exp(-4.0) * 4.0 ** k / factorial(k) for k in range(11)])
    assert np.abs(freq - pmf).max() < 0.005, (freq, pmf)
    b = gm.MetropolisHastings(gm.BinomialTarget(10, 0.3), gm.RandomWalkProposal(), np.full((Cn, 1), 5, np.int32), ctx=ctx).seed(42)
    out, st = b.run_progress(2000, 2000)
    assert out.min() >= 0 and out.max() <= 10
    freq = np.bincount(out.astype(np.int64).ravel(), minlength=11)[:11] / out.size
    pmf = np.array([comb(10, k) * 0.3 ** k * 0.7 ** (10 - k) for k in range(11)])
    assert np.abs(freq - pmf).max() < 0.005, (freq, pmf)
    assert st.rhat_std.max < 1.01
    # sharding and continuation invariance (Philox keyed by the global chain index)
    x0 = np.arange(300, dtype=np.int32).reshape(300, 1) % 7
    full = gm.MetropolisHastings(gm.PoissonTarget(4.0), gm.RandomWalkProposal(), x0, ctx=ctx).seed(9).run(30, 4)
    lo = gm.MetropolisHastings(gm.PoissonTarget(4.0), gm.RandomWalkProposal(), x0[:100], ctx=ctx, chain_offset=0).seed(9).run(30, 4)
    hi = gm.MetropolisHastings(gm.PoissonTarget(4.0), gm.RandomWalkProposal(), x0[100:], ctx=ctx, chain_offset=100).seed(9).run(30, 4)
    assert np.array_equal(full, np.concatenate([lo, hi]))
