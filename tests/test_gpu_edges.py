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
    for Cn, d in [(1, 7), (33, 5), (129, 37), (1000, 3)]:
        rng = np.random.default_rng(Cn * 100 + d)
        q0 = (1.0 + 0.2 * rng.standard_normal((Cn, d))).astype(np.float32)
        mom = rng.standard_normal((2, Cn, d)).astype(np.float32)
        ln_u = np.log(rng.random((2, Cn))).astype(np.float32)
        ref = oracle.hmc_run(oracle.ROSENBROCK_ND, [], q0, 0.01, 4, mom, ln_u)
        s = gm.HMC(gm.RosenbrockND(d), q0, 0.01, 4, seed=1, ctx=ctx).set_math_mode(True)
        s.inject(mom, ln_u)
        assert np.array_equal(s.run(2, 0), ref["samples"]), (Cn, d)


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
