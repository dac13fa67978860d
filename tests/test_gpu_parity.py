"""GPU parity tests: the CUDA path (through the C ABI, via general_mcmc_b200.api) against the CPU oracle
on identical injected randomness.  Run on a B200 with `pytest -m gpu`.

Bars (BASELINE.json north_star):
  * EXACT math mode, transcendental-free targets: bit-for-bit the oracle (positions, momenta, log
    accept, decisions) over whole multi-step runs.
  * FAST math mode: per-step equivalence within rel 1e-5 (f32) / 1e-10 (f64) over L leapfrog steps,
    accept decisions identical away from razor-thin margins (none in the seeded sets below).
  * Philox integer stream: bit-exact against the host restatement.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import general_mcmc_b200 as gm  # noqa: E402
from general_mcmc_b200 import _lib as L  # noqa: E402
import ctypes as C  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    return gm.default_context()


def _rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.max(np.abs(a - b) / (np.abs(b) + 1.0))


# -------------------------------------------------------------------------------------------------
def test_philox_device_bit_exact(ctx, oracle):
    rng = np.random.default_rng(0)
    ctr = rng.integers(0, 2**32, size=(257, 4), dtype=np.uint64).astype(np.uint32)
    key = np.array([0xDEADBEEF, 0x12345678], np.uint32)
    out = np.zeros_like(ctr)
    L.check(L.lib().gmcmc_philox_blocks(ctx._h, L.ptr(ctr), C.c_size_t(ctr.shape[0]), L.ptr(key), L.ptr(out)))
    ref = np.stack([oracle.philox4x32_10(c, key) for c in ctr])
    assert np.array_equal(out, ref)


def _dense_params(d, seed=0):
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.standard_normal((d, d)))
    lam = np.logspace(-1, 1, d)
    cov = (q * lam) @ q.T
    return gm.DenseGaussian(rng.standard_normal(d) * 0.1, cov=cov)


def _mixture(d, K=4):
    mu = np.stack([(k - 1.5) * (2.0 / np.sqrt(d)) * np.ones(d) for k in range(K)])
    return gm.GaussianMixture(np.full(K, 1.0 / K), mu, 1.0)


TARGETS_EXACT = [
    ("rosen3", lambda: gm.RosenbrockND(3), 3),
    ("rosen100", lambda: gm.RosenbrockND(100), 100),
    ("rosen37", lambda: gm.RosenbrockND(37), 37),
    ("iso5", lambda: gm.IsotropicGaussian(1.5, 5), 5),
    ("dense10", lambda: _dense_params(10), 10),
    ("dense70", lambda: _dense_params(70), 70),
    ("rosen2d", lambda: gm.Rosenbrock2D(1.0, 100.0), 2),
    ("dgauss2d", lambda: gm.DiffableGaussian2D([0.0, 1.0], [[4.0, 2.0], [2.0, 3.0]]), 2),
    ("gauss2d", lambda: gm.Gaussian2D([0.0, 1.0], [[4.0, 2.0], [2.0, 3.0]]), 2),
]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name,mk,d", TARGETS_EXACT, ids=[t[0] for t in TARGETS_EXACT])
def test_logp_grad_exact_mode_is_bit_exact(ctx, oracle, name, mk, d, dtype):
    tgt = mk()
    rng = np.random.default_rng(1)
    x = (rng.standard_normal((67, d)) * 0.7).astype(dtype)
    lp, g = tgt.logp_and_grad(x, ctx, exact=True)
    for i in range(x.shape[0]):
        rlp, rg = oracle.target_logp_grad(tgt.kind, d, tgt.params(), x[i])
        assert lp[i] == dtype(rlp), (name, i)
        assert np.array_equal(g[i], rg), (name, i)


@pytest.mark.parametrize("dtype,tol", [(np.float32, 2e-5), (np.float64, 1e-12)])
def test_logp_grad_fast_mode_and_mixture(ctx, oracle, dtype, tol):
    rng = np.random.default_rng(2)
    for tgt, d in [(gm.RosenbrockND(100), 100), (_dense_params(33), 33), (_mixture(100), 100), (_mixture(7, 3), 7)]:
        x = (rng.standard_normal((40, d)) * 0.5).astype(dtype)
        for exact in (False, True):
            lp, g = tgt.logp_and_grad(x, ctx, exact=exact)
            for i in range(x.shape[0]):
                rlp, rg = oracle.target_logp_grad(tgt.kind, d, tgt.params(), x[i])
                scale = np.abs(rg).max() + 1.0
                assert abs(lp[i] - rlp) <= tol * (abs(rlp) + 1.0) * 4
                assert np.max(np.abs(g[i] - rg)) <= tol * scale * 4


# -------------------------------------------------------------------------------------------------
# config 1: examples/rosenbrock3d_hmc — 4 chains, 3-D, eps 0.01, L 10, 50 discard + 400 collect
# -------------------------------------------------------------------------------------------------
CFG1_START = np.array([[0.30471708, -1.03998411, 0.7504512], [0.94056472, -1.95103519, -1.30217951],
                       [0.1278404, -0.31624259, -0.01680116], [-0.85304393, 0.87939797, 0.77779194]])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_cfg1_rosenbrock3d_hmc_exact_mode_bit_exact(ctx, oracle, dtype):
    n_discard, n_collect, Cn, d = 50, 400, 4, 3
    n = n_discard + n_collect
    rng = np.random.default_rng(42)
    mom = rng.standard_normal((n, Cn, d)).astype(dtype)
    ln_u = np.log(rng.random((n, Cn))).astype(dtype)
    q0 = CFG1_START.astype(dtype)
    ref = oracle.hmc_run(oracle.ROSENBROCK_ND, [], q0, 0.01, 10, mom, ln_u, want_traj=True)
    s = gm.HMC(gm.RosenbrockND(3), q0, 0.01, 10, seed=1, ctx=ctx).set_math_mode(True)
    s.inject(mom, ln_u)
    out = s.run(n_collect, n_discard)
    diag = s.diagnostics()
    assert np.array_equal(diag["accepted"], ref["accepted"])
    assert np.array_equal(diag["log_accept"], ref["log_accept"])
    assert np.array_equal(diag["prop_q"], ref["prop_q"])
    assert np.array_equal(diag["prop_p"], ref["prop_p"])
    assert np.array_equal(out, ref["samples"][:, n_discard:, :])
    assert np.array_equal(s.positions(), ref["q"])
    assert 0.3 < diag["accepted"].mean() <= 1.0
    # and against the committed golden fixture (tests/golden/oracle_cfg1.json, written by make_golden.py)
    import json
    import os
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_cfg1.json")))
    key = "f32" if dtype == np.float32 else "f64"
    assert np.array_equal(s.positions().astype(np.float64), np.asarray(gold[key]["final_positions"]))
    assert diag["accepted"].sum(0).tolist() == gold[key]["accepted_per_chain"]
    assert np.array_equal(out[:, -1, :].astype(np.float64), np.asarray(gold[key]["sample_399"]))


@pytest.mark.parametrize("dtype,tol", [(np.float32, 1e-5), (np.float64, 1e-10)])
@pytest.mark.parametrize("d,L,eps", [(3, 10, 0.01), (100, 32, 0.005), (37, 8, 0.01), (2, 10, 0.02)])
def test_hmc_per_step_equivalence_fast_mode(ctx, oracle, dtype, tol, d, L, eps):
    """4096 chains, one transition each from random starts: end-of-trajectory q, p, log accept within
    the north-star tolerance; decisions identical."""
    Cn = 4096
    rng = np.random.default_rng(7 + d)
    q0 = (1.0 + 0.3 * rng.standard_normal((Cn, d))).astype(dtype)
    mom = rng.standard_normal((1, Cn, d)).astype(dtype)
    ln_u = np.log(rng.random((1, Cn))).astype(dtype)
    ref = oracle.hmc_run(oracle.ROSENBROCK_ND, [], q0, eps, L, mom, ln_u, want_traj=True)
    s = gm.HMC(gm.RosenbrockND(d), q0, eps, L, seed=1, ctx=ctx)
    s.inject(mom, ln_u)
    out = s.run(1, 0)
    diag = s.diagnostics()
    assert _rel(diag["prop_q"], ref["prop_q"]) <= tol
    assert _rel(diag["prop_p"], ref["prop_p"]) <= tol * 50      # momenta scale with |grad| ~ 1e2..1e3
    # log accept is a difference of large energies: tolerance relative to each chain's own energy scale
    ke = 0.5 * (mom[0].astype(np.float64) ** 2).sum(-1) + 0.5 * (ref["prop_p"][0].astype(np.float64) ** 2).sum(-1)
    escale = (np.abs(ref["logp_cur"][0]) + np.abs(ref["logp_prop"][0]) + ke + 1.0)[None, :]
    err = np.abs(diag["log_accept"].astype(np.float64) - ref["log_accept"])
    assert np.all(err <= tol * escale * 8)
    margin = np.abs(ref["log_accept"].astype(np.float64) - ln_u)
    safe = margin > tol * escale * 8
    assert np.array_equal(diag["accepted"][safe], ref["accepted"][safe])
    assert (~safe).sum() <= Cn // 100
    assert out.shape == (Cn, 1, d)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mk,d,eps,L", [
    (lambda: gm.RosenbrockND(100), 100, 0.004, 16),
    (lambda: gm.RosenbrockND(37), 37, 0.01, 5),
    (lambda: _dense_params(24), 24, 0.1, 6),
    (lambda: gm.IsotropicGaussian(2.0, 9), 9, 0.5, 4),
    (lambda: gm.DiffableGaussian2D([0.0, 1.0], [[4.0, 2.0], [2.0, 3.0]]), 2, 0.1, 10),
    (lambda: gm.Rosenbrock2D(1.0, 100.0), 2, 0.01, 10),
], ids=["rosen100", "rosen37", "dense24", "iso9", "dgauss2d", "rosen2d"])
def test_hmc_multi_step_exact_mode_bit_exact(ctx, oracle, dtype, mk, d, eps, L):
    Cn, n = 333, 12
    tgt = mk()
    rng = np.random.default_rng(11)
    q0 = (1.0 + 0.2 * rng.standard_normal((Cn, d))).astype(dtype)
    mom = rng.standard_normal((n, Cn, d)).astype(dtype)
    ln_u = np.log(rng.random((n, Cn))).astype(dtype)
    ref = oracle.hmc_run(tgt.kind, tgt.params(), q0, eps, L, mom, ln_u, want_traj=True)
    s = gm.HMC(tgt, q0, eps, L, seed=3, ctx=ctx).set_math_mode(True)
    s.inject(mom, ln_u)
    out = s.run(n, 0)
    diag = s.diagnostics()
    assert np.array_equal(diag["accepted"], ref["accepted"])
    assert np.array_equal(diag["log_accept"], ref["log_accept"])
    assert np.array_equal(diag["prop_q"], ref["prop_q"])
    assert np.array_equal(out, ref["samples"])


# -------------------------------------------------------------------------------------------------
# Metropolis–Hastings
# -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mk,d", [
    (lambda: gm.Gaussian2D([0.0, 0.0], [[1.0, 0.0], [0.0, 1.0]]), 2),
    (lambda: gm.Gaussian2D([0.0, 1.0], [[4.0, 2.0], [2.0, 3.0]]), 2),
    (lambda: gm.IsotropicGaussian(1.3, 5), 5),
    (lambda: gm.RosenbrockND(3), 3),
], ids=["gauss2d_id", "gauss2d_cov", "iso5", "rosen3"])
def test_mh_exact_mode_bit_exact(ctx, oracle, dtype, mk, d):
    Cn, n, n_discard = 301, 45, 7
    tgt = mk()
    rng = np.random.default_rng(5)
    x0 = rng.standard_normal((Cn, d)).astype(dtype)
    z = rng.standard_normal((n, Cn, d)).astype(dtype)
    ln_u = np.log(rng.random((n, Cn))).astype(dtype)
    ref = oracle.mh_run(tgt.kind, tgt.params(), x0, 0.8, z, ln_u)
    s = gm.MetropolisHastings(tgt, gm.IsotropicGaussian(0.8), x0, ctx=ctx).set_math_mode(True)
    s.inject(z, ln_u)
    out = s.run(n - n_discard, n_discard)
    diag = s.diagnostics()
    assert out.dtype == np.float64
    assert np.array_equal(diag["accepted"], ref["accepted"])
    assert np.array_equal(diag["log_accept"], ref["log_ratio"])
    assert np.array_equal(out, ref["samples"][:, n_discard:, :])
    assert np.array_equal(s.positions(), ref["x"])


@pytest.mark.parametrize("dtype,tol", [(np.float32, 1e-5), (np.float64, 1e-12)])
def test_mh_fast_mode_matches_oracle(ctx, oracle, dtype, tol):
    Cn, n = 2000, 30
    tgt = gm.Gaussian2D([0.0, 1.0], [[4.0, 2.0], [2.0, 3.0]])
    rng = np.random.default_rng(6)
    x0 = rng.standard_normal((Cn, 2)).astype(dtype)
    z = rng.standard_normal((n, Cn, 2)).astype(dtype)
    ln_u = np.log(rng.random((n, Cn))).astype(dtype)
    ref = oracle.mh_run(tgt.kind, tgt.params(), x0, 1.0, z, ln_u)
    s = gm.MetropolisHastings(tgt, gm.IsotropicGaussian(1.0), x0, ctx=ctx)
    s.inject(z, ln_u)
    out = s.run(n, 0)
    diag = s.diagnostics()
    # decisions can only differ at razor-thin margins; the chains that never hit one must match fully
    margin = np.abs(ref["log_ratio"].astype(np.float64) - ln_u)
    thin = (margin < tol * 8 * (1.0 + np.abs(ref["log_ratio"]))).any(axis=0)
    assert thin.sum() <= Cn // 200
    ok = ~thin
    assert np.array_equal(diag["accepted"][:, ok], ref["accepted"][:, ok])
    assert np.allclose(out[ok], ref["samples"][ok], rtol=tol * 10, atol=tol * 10)


def test_mh_staging_tail_and_sharding_invariance(ctx):
    """Philox streams are keyed by (seed, global chain, transition): two shards with chain offsets
    reproduce the one-sampler run bit-for-bit, and run(a) + run(b) == run(a + b)."""
    Cn, n = 1000, 37     # 37 is not a multiple of the 16-step staging depth
    tgt = gm.Gaussian2D([0.0, 0.0], [[1.0, 0.0], [0.0, 1.0]])
    x0 = np.random.default_rng(8).standard_normal((Cn, 2))
    full = gm.MetropolisHastings(tgt, gm.IsotropicGaussian(1.0), x0, ctx=ctx).seed(42).run(n, 3)
    lo = gm.MetropolisHastings(tgt, gm.IsotropicGaussian(1.0), x0[:400], ctx=ctx, chain_offset=0).seed(42).run(n, 3)
    hi = gm.MetropolisHastings(tgt, gm.IsotropicGaussian(1.0), x0[400:], ctx=ctx, chain_offset=400).seed(42).run(n, 3)
    assert np.array_equal(full, np.concatenate([lo, hi]))
    s = gm.MetropolisHastings(tgt, gm.IsotropicGaussian(1.0), x0, ctx=ctx).seed(42)
    a = s.run(20, 3)
    b = s.run(17, 0)
    assert np.array_equal(full, np.concatenate([a, b], axis=1))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name,mk", [
    ("gauss2d", lambda: gm.Gaussian2D([0.5, -1.0], [[2.0, 0.6], [0.6, 1.0]])),
    ("dgauss2d", lambda: gm.DiffableGaussian2D([0.0, 1.0], [[4.0, 2.0], [2.0, 3.0]])),
    ("iso2", lambda: gm.IsotropicGaussian(1.5, 2)),
    ("rosen2d", lambda: gm.Rosenbrock2D(1.0, 100.0)),
    ("rosen_nd2", lambda: gm.RosenbrockND(2)),
])
def test_mh_2d_fast_kernel_agrees_with_exact_mode(ctx, name, mk, dtype):
    """The d = 2 fast kernel (mh_run2_kernel: its own draw mapping, MUFU Box-Muller, linear-domain accept test with
    full-precision fallback) against the exact-mode kernel (libdevice math, reference operation order) on every
    2-D target and both state types: different random streams, so the comparison is statistical — acceptance rate,
    per-coordinate medians and inter-quartile ranges over 8192 chains x 200 recorded steps."""
    Cn, n = 8192, 200
    x0 = (np.random.default_rng(3).standard_normal((Cn, 2)) * 0.5 + [0.5, 0.5]).astype(dtype)
    res = {}
    for exact in (False, True):
        s = gm.MetropolisHastings(mk(), gm.IsotropicGaussian(0.7), x0, ctx=ctx).seed(7).set_math_mode(exact)
        out = s.run(n, 300)
        assert out.dtype == np.float64 and np.isfinite(out).all()
        c = s.counters()
        flat = out.reshape(-1, 2)
        q = np.quantile(flat, [0.25, 0.5, 0.75], axis=0)
        res[exact] = (c.accepts / c.transitions, q[1], q[2] - q[0])
        moved = (np.diff(out, axis=1) != 0).any(axis=2).mean()
        assert abs(moved - res[exact][0]) < 0.02            # the recorded chain moves exactly when a proposal is accepted
    (a0, m0, w0), (a1, m1, w1) = res[False], res[True]
    assert abs(a0 - a1) < 0.01, (name, a0, a1)
    assert np.allclose(m0, m1, atol=0.03 * np.maximum(w1, 0.1)), (name, m0, m1)
    assert np.allclose(w0, w1, rtol=0.04), (name, w0, w1)


def test_hmc_sharding_and_continuation_invariance(ctx):
    Cn, d = 777, 100
    q0 = (1.0 + 0.1 * np.random.default_rng(9).standard_normal((Cn, d))).astype(np.float32)
    full = gm.HMC(gm.RosenbrockND(d), q0, 0.004, 8, seed=42, ctx=ctx).run(9, 2)
    lo = gm.HMC(gm.RosenbrockND(d), q0[:300], 0.004, 8, seed=42, ctx=ctx, chain_offset=0).run(9, 2)
    hi = gm.HMC(gm.RosenbrockND(d), q0[300:], 0.004, 8, seed=42, ctx=ctx, chain_offset=300).run(9, 2)
    assert np.array_equal(full, np.concatenate([lo, hi]))
    s = gm.HMC(gm.RosenbrockND(d), q0, 0.004, 8, seed=42, ctx=ctx)
    a = s.run(4, 2)
    b = s.run(5, 0)
    assert np.array_equal(full, np.concatenate([a, b], axis=1))


# -------------------------------------------------------------------------------------------------
# statistics
# -------------------------------------------------------------------------------------------------
def _ar1(c, n, p, rho, seed):
    rng = np.random.default_rng(seed)
    x = np.zeros((c, n, p), np.float32)
    e = rng.standard_normal((c, n, p)).astype(np.float32)
    x[:, 0] = e[:, 0]
    for t in range(1, n):
        x[:, t] = rho * x[:, t - 1] + np.sqrt(1 - rho * rho) * e[:, t]
    return x + np.arange(p, dtype=np.float32)[None, None, :]


@pytest.mark.parametrize("c,n,p,rho", [(4, 1000, 3, 0.0), (6, 1000, 13, 0.7), (3, 150, 8, 0.5), (5, 201, 1, 0.3),
                                        (64, 64, 20, 0.9), (2, 4, 2, 0.0), (16, 2001, 9, 0.95), (4, 5000, 2, 0.99),
                                        # warp-FFT path (128 <= N <= 1024): every padded length, ragged parameter blocks, odd n
                                        (7, 500, 20, 0.8), (5, 100, 9, 0.4), (3, 513, 5, 0.6), (300, 500, 17, 0.5), (9, 66, 8, 0.2)])
def test_split_rhat_ess_matches_oracle(ctx, oracle, c, n, p, rho):
    x = _ar1(c, n, p, rho, seed=c * 1000 + n)
    rhat, ess = gm.split_rhat_mean_ess(x, ctx)
    rrhat, ress = oracle.split_rhat_mean_ess(x)
    assert np.allclose(rhat, rrhat, rtol=2e-5, atol=1e-6)
    # ESS: two f32 FFT pipelines (per-chain inverse then mean, vs summed spectrum then one inverse)
    assert np.allclose(ess, ress, rtol=5e-4), (ess, ress)
    x64 = x.astype(np.float64)
    rhat64, ess64 = gm.split_rhat_mean_ess(x64, ctx)
    assert np.allclose(rhat64, rhat, rtol=1e-6) and np.allclose(ess64, ess, rtol=1e-6)


@pytest.mark.parametrize("c,n,p,dtype", [(2, 2, 2, np.float32), (5, 40, 3, np.float32), (64, 300, 17, np.float64),
                                          (6000, 9, 4, np.float32), (300, 500, 100, np.float32)])
def test_tracker_stats_match_oracle(ctx, oracle, c, n, p, dtype):
    """MultiChainTracker (stats.rs:199-339) on the device vs the oracle's host restatement (itself pinned by the
    reference's tracker R-hat KAT, stats.rs:734-783): chains with rejected steps (repeated rows), so the EMA
    acceptance rate is exercised; p_accept is bit-exact (same f32 fold), R-hat to f32 summation order."""
    rng = np.random.default_rng(c * 7 + n)
    x = _ar1(c, n, p, 0.6, seed=c + n).astype(dtype)
    stay = rng.random((c, n)) < 0.35
    for t in range(1, n):                      # a rejected proposal repeats the previous draw
        x[stay[:, t], t] = x[stay[:, t], t - 1]
    got = gm.tracker_stats(x, ctx)
    rhat, pa = oracle.tracker_rhat(np.ascontiguousarray(x.transpose(1, 0, 2)))
    assert got["p_accept"] == pa
    assert np.allclose(got["rhat"], rhat, rtol=2e-4 if c > 1000 else 2e-5)
    assert np.isclose(got["max_rhat"], rhat.max(), rtol=2e-4)


def test_run_stats_struct_matches_oracle(ctx, oracle):
    x = _ar1(8, 600, 11, 0.6, seed=3)
    st = gm.RunStats.from_samples(x, ctx)
    rrhat, ress = oracle.split_rhat_mean_ess(x)
    be, br = oracle.basic_stats(ress), oracle.basic_stats(rrhat)
    for k in ("min", "median", "max", "mean"):
        assert np.isclose(getattr(st.ess, k), be[k], rtol=5e-4)
        assert np.isclose(getattr(st.rhat, k), br[k], rtol=2e-5)
    assert np.isclose(st.rhat_std.max, 1.0 / br["min"], rtol=2e-5)


def test_ess_iid_uniform_reference_bands(ctx):
    """stats.rs:841-865: 4 x 1000 iid U(0,1): min ESS > 3800 and max R-hat < 1.01."""
    x = np.random.default_rng(0).random((4, 1000, 3)).astype(np.float32)
    rhat, ess = gm.split_rhat_mean_ess(x, ctx)
    assert ess.min() > 3800 * 0.9 and rhat.max() < 1.01


# -------------------------------------------------------------------------------------------------
# distributional agreement (Philox path)
# -------------------------------------------------------------------------------------------------
def test_hmc_gaussian2d_distribution_and_reference_ess_band(ctx):
    """hmc.rs:513-669: DiffableGaussian2D mu=[0,1], cov=[[4,2],[2,3]], eps 0.1, L 10: mean/cov within
    Monte-Carlo error, R-hat < 1.01 at many chains."""
    Cn = 1024
    tgt = gm.DiffableGaussian2D([0.0, 1.0], [[4.0, 2.0], [2.0, 3.0]])
    s = gm.HMC(tgt, np.zeros((Cn, 2), np.float32), 0.1, 10, seed=42, ctx=ctx)
    out, st = s.run_progress(1000, 500)
    per_chain_ess = st.ess.min / Cn
    assert 35 < per_chain_ess < 100          # reference band (3 chains x 1000 draws): ESS in [135, 230] => 45..77 per chain
    # split R-hat is ~ sqrt(1 + 1/ESS_per_split_chain): < 1.01 needs ~4000 draws per chain at this autocorrelation
    out, st = s.run_progress(6000, 0)
    flat = out.reshape(-1, 2).astype(np.float64)
    assert np.allclose(flat.mean(0), [0.0, 1.0], atol=0.02)
    assert np.allclose(np.cov(flat.T), [[4.0, 2.0], [2.0, 3.0]], atol=0.06)
    assert st.rhat_std.max < 1.01 and st.rhat.min > 0.99
    c = s.counters()
    assert 0.85 < c.accept_rate <= 1.0
    assert c.grad_evals == Cn * 7500 * 10


def test_mh_gaussian2d_distribution(ctx):
    """metropolis_hastings.rs:342-406 / tests/metrohast_2d_gaussian_test.rs: mean within 0.3, cov within 0.5
    (far tighter here with many chains)."""
    Cn = 1024
    tgt = gm.Gaussian2D([0.0, 1.0], [[4.0, 2.0], [2.0, 3.0]])
    s = gm.MetropolisHastings(tgt, gm.IsotropicGaussian(1.0), np.zeros((Cn, 2)), ctx=ctx).seed(42)
    out, st = s.run_progress(12000, 500)
    flat = out.reshape(-1, 2)
    assert np.allclose(flat.mean(0), [0.0, 1.0], atol=0.03)
    assert np.allclose(np.cov(flat.T), [[4.0, 2.0], [2.0, 3.0]], atol=0.08)
    assert st.rhat_std.max < 1.01


def test_pooled_and_per_chain_adaptation_reach_target_accept(ctx):
    Cn, d = 2048, 20
    q0 = (1.0 + 0.05 * np.random.default_rng(4).standard_normal((Cn, d))).astype(np.float32)
    for mode in ("pooled", "per_chain"):
        s = gm.HMC(gm.RosenbrockND(d), q0, 0.01, 8, seed=42, ctx=ctx).set_adaptation(mode, 0.8)
        s.run(0, 300)
        c0 = s.counters()
        s.run(50, 0)
        c1 = s.counters()
        rate = (c1.accepts - c0.accepts) / (50 * Cn)
        assert 0.6 < rate < 0.95, (mode, rate, c1.step_size)
        assert 1e-4 < c1.step_size < 1.0


# -------------------------------------------------------------------------------------------------
# config 2's production kernel (mh_run2_kernel) per step against the oracle
# -------------------------------------------------------------------------------------------------
def _philox4x32_10_np(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Random123), uint64 arithmetic; inputs broadcastable uint32 arrays."""
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
    mask = np.uint64(0xFFFFFFFF)
    c0, c1, c2, c3 = [np.asarray(c, np.uint64) & mask for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = int(k0), int(k1)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & mask, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & mask, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32)


def _mh2_host_draws(seed, chains, t_abs):
    """Host restatement of the 2-D fast kernel's draw contract (include/gmcmc.h, mh_kernel.cuh derive()): transition t
    takes words (0,1) [t even] or (2,3) [t odd] of Philox block (gchain, t >> 1, stream 0, block 0); 20-bit radius and
    angle grids, 23-bit accept uniform.  Returns z [T, C, 2] f64 (exact Box-Muller on the grid) and u [T, C] f64."""
    ch = np.asarray(chains, np.uint64)[None, :]
    t = np.asarray(t_abs, np.uint64)[:, None]
    r = _philox4x32_10_np(ch & np.uint64(0xFFFFFFFF), ch >> np.uint64(32), t >> np.uint64(1), np.uint64(0),
                          seed & 0xFFFFFFFF, seed >> 32)
    odd = (t & np.uint64(1)).astype(bool) & np.ones_like(ch, bool)
    w0 = np.where(odd, r[2], r[0]).astype(np.uint64)
    w1 = np.where(odd, r[3], r[1]).astype(np.uint64)
    u_rad = ((w0 >> np.uint64(12)).astype(np.float64) + 0.5) * 2.0 ** -20
    ang = 2.0 * np.pi * ((w1 >> np.uint64(12)).astype(np.float64) + 0.5) * 2.0 ** -20 - np.pi
    rad = np.sqrt(-2.0 * np.log(u_rad))
    z = np.stack([rad * np.cos(ang), rad * np.sin(ang)], axis=-1)
    k = ((w0 & np.uint64(0xFFF)) << np.uint64(11)) | (w1 & np.uint64(0x7FF))
    u = (k.astype(np.float64) + 0.5) * 2.0 ** -23
    return z, u


def test_philox_numpy_restatement_matches_oracle(oracle):
    rng = np.random.default_rng(0)
    ctr = rng.integers(0, 2**32, size=(64, 4), dtype=np.uint64).astype(np.uint32)
    key = np.array([0x2A, 0x7], np.uint32)
    got = np.stack(_philox4x32_10_np(ctr[:, 0], ctr[:, 1], ctr[:, 2], ctr[:, 3], key[0], key[1]), axis=1)
    assert np.array_equal(got, np.stack([oracle.philox4x32_10(c, key) for c in ctr]))


@pytest.mark.parametrize("name,mk,std", [
    ("gauss2d_id", lambda: gm.Gaussian2D([0.0, 0.0], [[1.0, 0.0], [0.0, 1.0]]), 1.0),       # BASELINE config 2
    ("gauss2d_cov", lambda: gm.Gaussian2D([0.0, 1.0], [[4.0, 2.0], [2.0, 3.0]]), 0.7),
])
def test_mh_2d_fast_kernel_per_step_vs_oracle(ctx, oracle, name, mk, std):
    """The kernel that carries the config-2 number draws from Philox only, so it cannot take injected draws.  Instead
    (i) the draws it USED (recorded by its instrumented instantiation, same arithmetic) are checked against a host
    restatement of its documented draw contract computed from bit-exact Philox words — the accept uniform bit for bit,
    the Box-Muller normals to the accuracy of the MUFU lg2 / sqrt / sin / cos approximations; (ii) those recorded draws
    are fed to the oracle's MHMarkovChain::step (metropolis_hastings.rs:306-318): decisions identical away from
    rounding-thin margins, states within 1e-6 over 1,000 steps x 4,096 chains; (iii) the instrumented launch and a
    plain launch of the production kernel produce bit-identical samples."""
    Cn, n, seed, off = 4096, 1000, 42, 1000
    tgt = mk()
    x0 = np.random.default_rng(12).standard_normal((Cn, 2))
    s = gm.MetropolisHastings(tgt, gm.IsotropicGaussian(std), x0, ctx=ctx, chain_offset=off).seed(seed)
    s.record(n)
    out = s.run(n, 0)
    diag = s.diagnostics()
    draws = s.draws()
    plain = gm.MetropolisHastings(tgt, gm.IsotropicGaussian(std), x0, ctx=ctx, chain_offset=off).seed(seed).run(n, 0)
    assert np.array_equal(out, plain)                                            # (iii)
    # (i) draw contract
    z_host, u_host = _mh2_host_draws(seed, off + np.arange(Cn), np.arange(n))
    assert np.array_equal(draws[..., 2].astype(np.float64), u_host)
    # MUFU accuracy: lg2.approx has an ABSOLUTE error of ~2^-22 (so the squared radius -2 ln u is off by up to ~4e-7 however
    # small it is: near u = 1 that is a large relative error of a tiny radius — and the same for z and -z, the proposal stays
    # symmetric), sin / cos.approx ~5e-7 absolute plus the f32 rounding of the angle
    zd = draws[..., :2].astype(np.float64)
    r2_dev, r2_host = (zd ** 2).sum(-1), (z_host ** 2).sum(-1)
    r2_err = np.abs(r2_dev - r2_host)
    cross = zd[..., 0] * z_host[..., 1] - zd[..., 1] * z_host[..., 0]
    dot = (zd * z_host).sum(-1)
    far = r2_host > 1e-4
    ang_err = np.abs(np.arctan2(cross[far], dot[far]))
    print("mh2 %s: max |r2_dev - r2_host| %.2e (rel part %.2e), max angle error %.2e rad, max |dz| %.2e" % (
        name, r2_err.max(), (r2_err / (1.0 + r2_host)).max(), ang_err.max(), np.abs(zd - z_host).max()))
    assert np.all(r2_err <= 1e-6 + 3e-6 * r2_host)
    assert ang_err.max() < 4e-6
    assert np.abs(zd - z_host).max() < 2e-3        # worst case: the radius grid point next to u = 1
    # (ii) per-step parity given the draws
    zf = np.ascontiguousarray(draws[..., :2].astype(np.float64))
    ln_u = np.log(draws[..., 2].astype(np.float64))
    ref = oracle.mh_run(tgt.kind, tgt.params(), x0, std, zf, ln_u)
    margin = np.abs(ref["log_ratio"] - ln_u)
    thin = (margin < 1e-12 * (1.0 + np.abs(ref["log_ratio"]) + np.abs(ln_u))).any(axis=0)
    ok = ~thin
    print("mh2 %s: chains with a rounding-thin margin %d / %d, accept rate %.3f" % (name, thin.sum(), Cn, ref["accepted"].mean()))
    assert thin.sum() <= 2
    assert np.array_equal(diag["accepted"][:, ok], ref["accepted"][:, ok])
    assert np.abs(diag["log_accept"][:, ok] - ref["log_ratio"][:, ok]).max() < 1e-11
    assert np.abs(out[ok] - ref["samples"][ok]).max() < 1e-6
    assert np.abs(s.positions()[ok] - ref["x"][ok]).max() < 1e-6
    assert 0.2 < ref["accepted"].mean() < 0.8
