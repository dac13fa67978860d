"""Gibbs sampler (/root/reference/src/gibbs.rs): the reference's own test cases on the device path, per-sweep
parity against the oracle with injected draws, and an ahead-of-time compiled user conditional."""
import os

import numpy as np
import pytest

import oracle_lib as orc

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def gm():
    import general_mcmc_b200 as g
    return g


def test_gibbs_constant_conditional_like_reference():
    """gibbs.rs:248-289: a conditional that always returns c; run(10, 5) of 4 chains is [4, 10, 2] filled with c."""
    g = gm()
    s = g.GibbsSampler(g.ConstantConditional(42.0), g.init_det(4, 2, "float64")).set_seed(42)
    out = s.run(10, 5)
    assert out.shape == (4, 10, 2) and out.dtype == np.float64
    assert np.array_equal(out, np.full((4, 10, 2), 42.0))
    out2, stats = g.GibbsSampler(g.ConstantConditional(7.0), np.zeros((3, 3))).run_progress(10, 5)
    assert np.array_equal(out2, np.full((3, 10, 3), 7.0))
    s3 = g.GibbsSampler(g.ConstantConditional(13.0), np.array([[1.0, 2.0, 3.0]]))
    assert np.array_equal(s3.positions(), [[1.0, 2.0, 3.0]])     # gibbs.rs:461-488: the state before any step
    s3.step()
    assert np.array_equal(s3.positions(), [[13.0, 13.0, 13.0]])  # gibbs.rs:248-262


@pytest.mark.parametrize("mu0,sigma0,mu1,sigma1,pi0", [(-2.0, 1.0, 3.0, 1.5, 0.5), (-42.0, 69.0, 1.0, 2.0, 0.123)])
def test_gibbs_mixture_moments_like_reference(mu0, sigma0, mu1, sigma1, pi0):
    """gibbs.rs:291-383 (assert_mixture_simulation): marginal mean and variance of x within 10 % of the mixture's.  The
    reference runs 4 chains x 100,000 draws; here 4,096 chains x 2,000 after 500 sweeps from z = 0 / 1 starts."""
    g = gm()
    theo_mean = pi0 * mu0 + (1 - pi0) * mu1
    theo_var = pi0 * (sigma0 ** 2 + (mu0 - theo_mean) ** 2) + (1 - pi0) * (sigma1 ** 2 + (mu1 - theo_mean) ** 2)
    x0 = np.zeros((4096, 2))
    x0[::2, 1] = 1.0
    s = g.GibbsSampler(g.MixtureConditional(mu0, sigma0, mu1, sigma1, pi0), x0).set_seed(42)
    out = s.run(2000, 500)
    x, z = out[:, :, 0].ravel(), out[:, :, 1].ravel()
    assert set(np.unique(z)) <= {0.0, 1.0}
    assert abs(x.mean() - theo_mean) < abs(theo_mean) / 10
    assert abs(x.var(ddof=1) - theo_var) < theo_var / 10
    assert abs(z.mean() - (1 - pi0)) < 0.02
    # chains are independent (unlike the reference's cloned RNG state, gibbs.rs:145-148)
    assert not np.array_equal(out[0], out[2])


def test_gibbs_mixture_per_sweep_vs_oracle():
    """Injected normals / uniforms: the device sweep reproduces the oracle's GibbsMarkovChain::step bit for bit (x = mu +
    sigma * noise is exact arithmetic; z flips only where u is within an ulp of prob_z1)."""
    g = gm()
    rng = np.random.default_rng(5)
    Cn, n = 512, 300
    params = [-2.0, 1.0, 3.0, 1.5, 0.35]
    x0 = np.stack([rng.standard_normal(Cn), (rng.random(Cn) < 0.5).astype(np.float64)], axis=1)
    normals = rng.standard_normal((n, Cn, 2))
    unif = rng.random((n, Cn, 2))
    ref = orc.gibbs_run(1, params, x0, normals, unif)
    s = g.GibbsSampler(g.MixtureConditional(*params), x0).set_seed(1)
    s.inject(normals, unif)
    out = s.run(n, 0)
    same = (out == ref["samples"]).all(axis=(1, 2))
    assert same.mean() > 0.995, same.mean()
    assert np.array_equal(out[same], ref["samples"][same])
    assert np.array_equal(s.positions()[same], ref["x"][same])
    # burn-in + resumed runs: run(a) then run(b) == run(a + b)
    s2 = g.GibbsSampler(g.MixtureConditional(*params), x0).set_seed(9)
    a = s2.run(40, 10)
    b = s2.run(30, 0)
    s3 = g.GibbsSampler(g.MixtureConditional(*params), x0).set_seed(9)
    full = s3.run(80, 0)
    assert np.array_equal(a, full[:, 10:50]) and np.array_equal(b, full[:, 50:80])


def test_gibbs_sharded_chains_equal_one_sampler():
    g = gm()
    params = [0.0, 1.0, 4.0, 0.5, 0.4]
    x0 = np.zeros((256, 2))
    whole = g.GibbsSampler(g.MixtureConditional(*params), x0).set_seed(3).run(50, 5)
    lo = g.GibbsSampler(g.MixtureConditional(*params), x0[:100], chain_offset=0).set_seed(3).run(50, 5)
    hi = g.GibbsSampler(g.MixtureConditional(*params), x0[100:], chain_offset=100).set_seed(3).run(50, 5)
    assert np.array_equal(whole, np.concatenate([lo, hi]))


def test_gibbs_custom_conditional_plugin():
    """A user conditional compiled ahead of time (gmcmc_custom_conditional.cuh): bivariate normal with correlation rho."""
    g = gm()
    so = g.build_custom_conditional(os.path.join(HERE, "plugins", "bivariate_cond.cu"), os.path.join(HERE, "plugins", "bivariate_cond.so"))
    rho = 0.8
    s = g.GibbsSampler(g.CustomConditional(so, [rho]), np.zeros((2048, 2))).set_seed(11)
    out, stats = s.run_progress(1000, 100)
    flat = out.reshape(-1, 2)
    assert abs(flat.mean(axis=0)).max() < 0.02
    assert abs(flat.var(axis=0) - 1.0).max() < 0.03
    assert abs(np.corrcoef(flat.T)[0, 1] - rho) < 0.01
    assert stats.rhat.max < 1.01


def test_gibbs_argument_errors():
    g = gm()
    with pytest.raises(g.GmcmcError):
        g.GibbsSampler(g.MixtureConditional(0, 1, 1, 1, 0.5), np.zeros((4, 3)))      # state must be [x, z]
    with pytest.raises(g.GmcmcError):
        g.GibbsSampler(g.MixtureConditional(0, -1, 1, 1, 0.5), np.zeros((4, 2)))
    with pytest.raises(g.GmcmcError):
        g.GibbsSampler(g.CustomConditional("/nonexistent/libcond.so"), np.zeros((4, 2)))
