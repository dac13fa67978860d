"""Ahead-of-time compiled custom targets (north star: "custom targets plug in as AOT-compiled device logp/grad
functions registered through the same C ABI"): tests/plugins/banana.cu is compiled with nvcc against
csrc/gmcmc_custom_target.cuh and driven through gmcmc_target_create_custom."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import general_mcmc_b200 as gm  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "plugins", "banana.cu")
SO = os.path.join(HERE, "plugins", "banana.so")


@pytest.fixture(scope="module")
def banana():
    csrc = os.path.join(os.path.dirname(HERE), "general_mcmc_b200", "csrc")
    newest = max([os.path.getmtime(SRC)] + [os.path.getmtime(os.path.join(csrc, f)) for f in os.listdir(csrc)
                                            if f.endswith((".h", ".cuh", ".inc"))])
    if not os.path.exists(SO) or os.path.getmtime(SO) < newest:   # the plugin embeds the kernels: any header change rebuilds it
        gm.build_custom_target(SRC, SO)
    return gm.CustomTarget(SO, 3, [1.5, 0.3])


def _ref(x, s=1.5, b=0.3):
    r = x[:, 1] - b * x[:, 0] ** 2
    u = x[:, 2] - 1.0
    lp = -0.5 * x[:, 0] ** 2 / s**2 - 0.5 * r * r - 2.0 * u * u
    g = np.stack([-x[:, 0] / s**2 + 2 * b * x[:, 0] * r, -r, -4.0 * u], axis=1)
    return lp, g


@pytest.mark.parametrize("dtype,tol", [(np.float32, 1e-5), (np.float64, 1e-13)])
def test_custom_target_logp_grad(banana, dtype, tol):
    x = np.random.default_rng(0).standard_normal((513, 3)).astype(dtype)
    lp, g = banana.logp_and_grad(x)
    rlp, rg = _ref(x.astype(np.float64))
    assert np.allclose(lp, rlp, rtol=tol, atol=tol * 10)
    assert np.allclose(g, rg, rtol=tol, atol=tol * 10)


def test_custom_target_hmc_and_nuts_sample_it(banana):
    Cn = 4096
    q0 = np.zeros((Cn, 3), np.float32)
    q0[:, 2] = 1.0
    s = gm.HMC(banana, q0, 0.15, 10, seed=42)
    out, st = s.run_progress(600, 300)
    flat = out[:, ::5].reshape(-1, 3).astype(np.float64)
    assert abs(flat[:, 0].var() - 2.25) < 0.08          # x0 ~ N(0, 1.5^2)
    assert abs(flat[:, 1].mean() - 0.3 * 2.25) < 0.04   # E[x1] = b s^2
    assert abs(flat[:, 2].mean() - 1.0) < 0.01 and abs(flat[:, 2].var() - 0.25) < 0.01
    assert s.counters().accept_rate > 0.8
    n = gm.NUTS(banana, q0[:1024], 0.8, seed=7, max_depth=8)
    o2 = n.run(300, 200)
    f2 = o2.reshape(-1, 3).astype(np.float64)
    assert np.isfinite(f2).all()
    assert abs(f2[:, 0].var() - 2.25) < 0.2 and abs(f2[:, 2].mean() - 1.0) < 0.03


@pytest.mark.parametrize("dtype,exact", [(np.float64, False), (np.float32, False), (np.float64, True)])
def test_custom_target_metropolis_hastings(banana, dtype, exact):
    """MetropolisHastings::new(target, proposal, ..) with a custom Target (distributions.rs:107-110): the K2 kernel
    instantiated by the plugin for its log density.  Moments of the banana; per-step equivalence against a numpy
    restatement of metropolis_hastings.rs:306-318 on injected draws."""
    Cn = 4096
    x0 = np.zeros((Cn, 3), dtype)
    x0[:, 2] = 1.0
    s = gm.MetropolisHastings(banana, gm.IsotropicGaussian(0.6), x0).seed(42).set_math_mode(exact)
    out = s.run(1500, 500)
    assert out.dtype == np.float64 and out.shape == (Cn, 1500, 3)
    flat = out[:, ::10].reshape(-1, 3)
    assert abs(flat[:, 0].var() - 2.25) < 0.12          # x0 ~ N(0, 1.5^2)
    assert abs(flat[:, 1].mean() - 0.3 * 2.25) < 0.06   # E[x1] = b s^2
    assert abs(flat[:, 2].mean() - 1.0) < 0.02 and abs(flat[:, 2].var() - 0.25) < 0.02
    c = s.counters()
    assert 0.2 < c.accepts / c.transitions < 0.7
    # injected draws: the accept decisions and states follow the reference recurrence
    Ci, n = 64, 12
    rng = np.random.default_rng(1)
    xi = rng.standard_normal((Ci, 3)).astype(dtype)
    z = rng.standard_normal((n, Ci, 3)).astype(dtype)
    lnu = np.log(rng.random((n, Ci))).astype(dtype)
    t = gm.MetropolisHastings(banana, gm.IsotropicGaussian(0.6), xi).seed(1).set_math_mode(exact)
    t.inject(z, lnu)
    got = t.run(n, 0)
    x = xi.astype(np.float64)
    alive = np.ones(Ci, bool)                  # chains none of whose decisions sat on the rounding margin so far
    tol = 1e-12 if dtype == np.float64 else 2e-5
    for k in range(n):
        xp = x + z[k].astype(np.float64) * 0.6
        lr = _ref(xp)[0] - _ref(x)[0]
        alive &= np.abs(lr - lnu[k]) > 1e-3
        x = np.where((lr > lnu[k])[:, None], xp, x)
        assert np.allclose(got[alive, k], x[alive], rtol=tol, atol=tol)
    assert alive.sum() > Ci // 2
