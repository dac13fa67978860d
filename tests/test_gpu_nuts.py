"""GPU parity tests of the NUTS kernel (K5) against the CPU oracle's recursive restatement of
generic_nuts.rs, on identical injected per-chain streams (normals, Exp(1), uniforms)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import general_mcmc_b200 as gm  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    return gm.default_context()


def _streams(Cn, d, steps, seed, n_unif=4000):
    rng = np.random.default_rng(seed)
    normals = rng.standard_normal((Cn, d * (steps + 2)))
    exp1 = rng.exponential(size=(Cn, steps + 2))
    unif = rng.random((Cn, n_unif))
    return normals, exp1, unif


def _dense(d, seed=0):
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.standard_normal((d, d)))
    lam = np.logspace(-0.5, 1.5, d)
    return gm.DenseGaussian(np.zeros(d), cov=(q * lam) @ q.T)


CASES = [
    ("iso5_std10", lambda: gm.IsotropicGaussian(10.0, 5), 5, 0.7, 0),
    ("iso5_depth3", lambda: gm.IsotropicGaussian(10.0, 5), 5, 0.7, 3),
    ("iso37", lambda: gm.IsotropicGaussian(6.0, 37), 37, 0.9, 0),
    ("dgauss2d", lambda: gm.DiffableGaussian2D([0.0, 1.0], [[40.0, 20.0], [20.0, 30.0]]), 2, 0.5, 0),
    ("dense24", lambda: _dense(24), 24, 0.6, 0),
    ("iso100", lambda: gm.IsotropicGaussian(8.0, 100), 100, 0.8, 6),
]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name,mk,d,eps0,max_depth", CASES, ids=[c[0] for c in CASES])
def test_nuts_exact_mode_bit_exact_no_adaptation(ctx, oracle, name, mk, d, eps0, max_depth, dtype):
    """n_discard = 0: the step size is eps0 for the first transition and eps_bar = 1 afterwards (the
    reference's behaviour, generic_nuts.rs:921-923), so no transcendental enters the trajectory: samples,
    leapfrog counts and consumed draws must be bit-identical to the recursive oracle."""
    Cn, n_collect = 48, 7
    tgt = mk()
    rng = np.random.default_rng(3)
    q0 = (rng.standard_normal((Cn, d)) * 3.0).astype(dtype)
    normals, exp1, unif = _streams(Cn, d, n_collect, seed=5)
    ref = oracle.nuts_run(tgt.kind, tgt.params(), q0, 0.8, max_depth, eps0, n_collect, 0, normals, exp1, unif)
    assert not ref["exhausted"].any()
    s = gm.NUTS(tgt, q0, 0.8, seed=1, ctx=ctx, max_depth=max_depth, init_step_size=eps0).set_math_mode(True)
    s.inject_streams(normals, exp1, unif)
    out = s.run(n_collect, 0)
    st = s.state()
    assert np.array_equal(st["leapfrogs"], ref["leapfrogs"]), name
    assert np.array_equal(st["used"].astype(np.int64), ref["used"]), name
    assert np.array_equal(out, ref["samples"]), name
    assert np.array_equal(out[:, 0, :], q0)            # run(): sample 0 is the initial position (nuts.rs:588-601)
    assert ref["leapfrogs"].mean() > 10                # the trees are non-trivial


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_nuts_find_reasonable_epsilon_bit_exact(ctx, oracle, dtype):
    """run(1, 0) takes no transition: the returned sample is the initial point (nuts.rs:588-601) and the
    per-chain step size is find_reasonable_epsilon's (generic_nuts.rs:1025-1102; reference KAT: exactly 2.0
    for N(0, I) at q = [0, 1], p = [1, 0], nuts.rs:508-519)."""
    tgt = gm.IsotropicGaussian(1.0, 2)
    q0 = np.array([[0.0, 1.0]], dtype)
    normals = np.array([[1.0, 0.0, 0.0, 0.0]])
    s = gm.NUTS(tgt, q0, 0.8, seed=1, ctx=ctx).set_math_mode(True)
    s.inject_streams(normals, np.ones((1, 2)), np.full((1, 8), 0.3))
    out = s.run(1, 0)
    assert np.array_equal(out[:, 0, :], q0)
    assert s.state()["eps"][0] == dtype(2.0)
    # many chains, several targets
    for tgt, d in [(gm.IsotropicGaussian(3.0, 17), 17), (gm.RosenbrockND(10), 10), (_dense(24), 24)]:
        Cn = 96
        rng = np.random.default_rng(9)
        q0 = (1.0 + 0.5 * rng.standard_normal((Cn, d))).astype(dtype)
        normals, exp1, unif = _streams(Cn, d, 1, seed=11, n_unif=8)
        ref = oracle.nuts_run(tgt.kind, tgt.params(), q0, 0.8, 0, -1.0, 1, 0, normals, exp1, unif)
        s = gm.NUTS(tgt, q0, 0.8, seed=1, ctx=ctx).set_math_mode(True)
        s.inject_streams(normals, exp1, unif)
        s.run(1, 0)
        assert np.array_equal(s.state()["eps"], ref["eps"])


@pytest.mark.parametrize("exact", [True, False])
def test_nuts_with_adaptation_matches_oracle(ctx, oracle, exact):
    """With warm-up the step size goes through exp / log / pow / sqrt (device libm vs glibc): the first
    transition is still bit-comparable, later ones agree for the chains whose tree decisions do not sit on a
    rounding boundary."""
    Cn, d, n_collect, n_discard = 64, 6, 4, 6
    tgt = gm.IsotropicGaussian(2.0, d)
    rng = np.random.default_rng(13)
    q0 = rng.standard_normal((Cn, d))
    normals, exp1, unif = _streams(Cn, d, n_collect + n_discard, seed=17)
    ref = oracle.nuts_run(tgt.kind, tgt.params(), q0, 0.8, 8, -1.0, n_collect, n_discard, normals, exp1, unif)
    s = gm.NUTS(tgt, q0, 0.8, seed=1, ctx=ctx, max_depth=8).set_math_mode(exact)
    s.inject_streams(normals, exp1, unif)
    out = s.run(n_collect, n_discard)
    st = s.state()
    same = (st["leapfrogs"] == ref["leapfrogs"])
    assert same.mean() > 0.9
    assert np.allclose(out[same], ref["samples"][same], rtol=1e-6, atol=1e-6)
    assert np.allclose(st["eps"][same], ref["eps"][same], rtol=1e-6)


def test_nuts_distribution_gaussian2d_and_rosenbrock_smoke(ctx):
    """Philox path: NUTS on the 2-D Gaussian of the reference's tests recovers mean / covariance; the
    Rosenbrock run stays finite and adapts to a sane step size."""
    Cn = 2048
    tgt = gm.DiffableGaussian2D([0.0, 1.0], [[4.0, 2.0], [2.0, 3.0]])
    s = gm.NUTS(tgt, np.zeros((Cn, 2), np.float32), 0.8, seed=42, ctx=ctx, max_depth=10)
    out, st = s.run_progress(400, 200)
    flat = out.reshape(-1, 2).astype(np.float64)
    assert np.isfinite(flat).all()
    assert np.allclose(flat.mean(0), [0.0, 1.0], atol=0.03)
    assert np.allclose(np.cov(flat.T), [[4.0, 2.0], [2.0, 3.0]], atol=0.12)
    assert st.rhat_std.max < 1.02
    c = s.counters()
    assert 0.05 < c.step_size < 3.0 and c.grad_evals > Cn * 600

    d = 10
    q0 = (1.0 + 0.1 * np.random.default_rng(1).standard_normal((512, d))).astype(np.float32)
    r = gm.NUTS(gm.RosenbrockND(d), q0, 0.8, seed=7, ctx=ctx, max_depth=8)
    out = r.run(50, 100)
    assert np.isfinite(out).all()
    assert 1e-4 < r.counters().step_size < 1.0


def test_nuts_sharding_invariance(ctx):
    Cn, d = 200, 12
    tgt = gm.IsotropicGaussian(1.5, d)
    q0 = np.random.default_rng(2).standard_normal((Cn, d)).astype(np.float32)
    full = gm.NUTS(tgt, q0, 0.8, seed=42, ctx=ctx, max_depth=6).run(6, 4)
    lo = gm.NUTS(tgt, q0[:72], 0.8, seed=42, ctx=ctx, max_depth=6, chain_offset=0).run(6, 4)
    hi = gm.NUTS(tgt, q0[72:], 0.8, seed=42, ctx=ctx, max_depth=6, chain_offset=72).run(6, 4)
    assert np.array_equal(full, np.concatenate([lo, hi]))


@pytest.mark.parametrize("dtype,exact,n_discard,cfg,n_upd", [
    (np.float64, True, 20, (1, 1, 10, 0.05, 1e-6), 2),
    (np.float64, True, 40, (3, 2, 10, 0.05, 1e-6), 3),
    (np.float32, True, 20, (1, 1, 10, 0.05, 1e-6), 2),
    (np.float32, False, 20, (1, 1, 10, 0.05, 1e-6), 2),
])
def test_nuts_diagonal_mass_adaptation_matches_oracle(ctx, oracle, dtype, exact, n_discard, cfg, n_upd):
    """GenericNUTS::new_with_mass_matrix (generic_nuts.rs:379-398), diagonal adaptation.  Window ends inside the
    warm-up (m = 11, 18 for start_buffer 1 / end_buffer 1 / window 10 / 20 warm-up transitions; m = 13, 23, 37 for
    3 / 2 / 10 / 40) each replace the mass matrix by the regularised running variance, probe a new step size from a
    fresh momentum and restart dual averaging.  Same injected streams on both sides.  The step size goes through
    exp / log / pow (device libm vs glibc, 1 ulp), and dual averaging feeds every rounding difference back into the
    next trajectory: measured deviations grow about tenfold every few transitions with or without the mass matrix
    (tools/diag_mass.py), so the bounds are on the per-chain MEDIAN error, with a loose cap on the worst chain."""
    Cn, d, n_collect = 96, 5, 6
    scales = np.array([0.3, 1.0, 3.0, 0.7, 2.0])
    tgt = gm.DenseGaussian(np.zeros(d), cov=np.diag(scales ** 2))
    rng = np.random.default_rng(21)
    q0 = (rng.standard_normal((Cn, d)) * scales).astype(dtype)
    normals, exp1, unif = _streams(Cn, d, n_collect + n_discard + 4, seed=23, n_unif=20000)
    ref = oracle.nuts_run(tgt.kind, tgt.params(), q0, 0.8, 8, -1.0, n_collect, n_discard, normals, exp1, unif, mass_cfg=cfg)
    assert not ref["exhausted"].any()
    s = gm.NUTS(tgt, q0, 0.8, seed=1, ctx=ctx, max_depth=8,
                mass_matrix=gm.NUTSMassMatrixConfig("diagonal", *cfg)).set_math_mode(exact)
    s.inject_streams(normals, exp1, unif)
    out = s.run(n_collect, n_discard)
    st = s.state()
    inv, n_updates = s.mass_matrix()
    assert n_updates == n_upd
    # normals consumed: d at init, d per transition, d per mass-matrix probe — on every chain, whatever its trees did
    assert np.array_equal(st["used"][:, 0].astype(np.int64), ref["used"][:, 0])
    assert (ref["used"][:, 0] == d * (1 + n_collect + n_discard - 1 + n_upd)).all()
    same = (st["leapfrogs"] == ref["leapfrogs"])
    assert same.mean() > (0.9 if dtype == np.float64 else 0.6)
    err = (np.abs(inv[same] - ref["mass_inv"][same]) / ref["mass_inv"][same]).max(1)
    med_tol, max_tol = (1e-8, 1e-3) if dtype == np.float64 else (2e-3, 0.5)
    assert np.median(err) < med_tol and err.max() < max_tol
    eps_err = np.abs(st["eps"][same] - ref["eps"][same]) / ref["eps"][same]
    assert np.median(eps_err) < med_tol and eps_err.max() < max_tol
    samp_err = np.abs(out[same] - ref["samples"][same]).reshape(same.sum(), -1).max(1)
    assert np.median(samp_err) < 100 * med_tol
    assert (inv > 0).all() and not np.allclose(inv, 1.0)       # the mass matrix did move away from the identity


def test_nuts_mass_adaptation_philox_matches_oracle_distribution(ctx, oracle):
    """Philox path with the reference's default windows (start 75 / end 50 / window 25), 300 warm-up transitions,
    chains started in stationarity on an ill-conditioned diagonal Gaussian.  The identity-mass sampler keeps the
    target variances.  The adapting sampler reproduces what the REFERENCE algorithm does, which is not what one would
    hope for: the reference sets the mass to the variance (generic_nuts.rs:196-206: inv = 1 / var, p = z sqrt(var)),
    searches the step size and tests sub-tree U-turns with the identity mass (:1009-1023, :1316) and the whole-tree
    U-turn with the adapted one (:872), so its transition is no longer reversible and the large-scale coordinates
    come out over-dispersed (CPU oracle: variance ratio ~1.9 on the widest coordinate).  Parity is the bar: the
    GPU must show the oracle's distribution, bias included."""
    Cn, d = 1024, 8
    scales = np.logspace(-1, 1, d)
    tgt = gm.DenseGaussian(np.zeros(d), cov=np.diag(scales ** 2))
    q0 = (np.random.default_rng(5).standard_normal((Cn, d)) * scales).astype(np.float32)
    r2 = np.random.default_rng(1)
    Co = 512
    streams = (r2.standard_normal((Co, d * 412)), r2.exponential(size=(Co, 402)), r2.random((Co, 400000)))
    for name, cfg in [("identity", None), ("diagonal", (75, 50, 25, 0.05, 1e-6))]:
        ref = oracle.nuts_run(tgt.kind, tgt.params(), q0[:Co], 0.8, 10, -1.0, 100, 300, *streams, mass_cfg=cfg)
        assert not ref["exhausted"].any()
        ratio_ref = ref["samples"].reshape(-1, d).astype(np.float64).var(0) / scales ** 2
        mm = gm.NUTSMassMatrixConfig("diagonal", *cfg) if cfg else None
        s = gm.NUTS(tgt, q0, 0.8, seed=11, ctx=ctx, max_depth=10, mass_matrix=mm)
        out = s.run(100, 300)
        ratio = out.reshape(-1, d).astype(np.float64).var(0) / scales ** 2
        leap_gpu = s.counters().grad_evals / (Cn * 399.0)
        leap_ref = ref["leapfrogs"].mean() / 399.0
        assert np.allclose(ratio, ratio_ref, rtol=0.15), (name, ratio, ratio_ref)
        assert abs(leap_gpu / leap_ref - 1.0) < 0.15, (name, leap_gpu, leap_ref)
        if cfg is None:
            assert np.allclose(ratio, 1.0, atol=0.06)
        else:
            assert ratio[-1] > 1.4 and ratio_ref[-1] > 1.4          # the reference's over-dispersion, reproduced
            inv, n = s.mass_matrix()
            assert n == 4 and np.all(np.isfinite(inv))
            # mass = regularised within-window variance: same per-coordinate medians over chains as the oracle's
            got = np.median(1.0 / inv.astype(np.float64), axis=0)
            want = np.median(1.0 / ref["mass_inv"].astype(np.float64), axis=0)
            assert np.allclose(got, want, rtol=0.25), (got, want)


def _corr_gauss(d, seed=4):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((d, d))
    cov = a @ a.T / d + np.diag(np.linspace(0.3, 2.0, d))
    return gm.DenseGaussian(np.zeros(d), cov=cov), cov


@pytest.mark.parametrize("dtype,exact,n_discard,cfg,n_upd", [
    (np.float64, True, 20, (1, 1, 10, 0.05, 1e-6), 2),
    (np.float64, True, 40, (3, 2, 10, 0.05, 1e-6), 3),
    (np.float32, True, 20, (1, 1, 10, 0.05, 1e-6), 2),
    (np.float32, False, 20, (1, 1, 10, 0.05, 1e-6), 2),
])
def test_nuts_dense_mass_adaptation_matches_oracle(ctx, oracle, dtype, exact, n_discard, cfg, n_upd):
    """MassMatrixAdaptation::Dense (generic_nuts.rs:36-39): running outer-product sums (RunningCov :81-132), regularised
    covariance -> Cholesky -> inverse per chain at every window end (dense_from_cov :208-226, :306-359, :970-997), momenta
    chol z, velocities / kinetic energy through the dense inverse, identity mass in the step-size search and the sub-tree
    U-turn tests.  Same injected streams on both sides; bounds as for the diagonal case (per-chain medians, see there)."""
    Cn, d, n_collect = 96, 5, 6
    tgt, cov = _corr_gauss(d)
    rng = np.random.default_rng(21)
    q0 = (rng.standard_normal((Cn, d)) @ np.linalg.cholesky(cov).T).astype(dtype)
    normals, exp1, unif = _streams(Cn, d, n_collect + n_discard + 4, seed=23, n_unif=20000)
    ref = oracle.nuts_run(tgt.kind, tgt.params(), q0, 0.8, 8, -1.0, n_collect, n_discard, normals, exp1, unif,
                          mass_cfg=cfg + (1, 75))
    assert not ref["exhausted"].any() and (ref["mass_updates"] == n_upd).all()
    s = gm.NUTS(tgt, q0, 0.8, seed=1, ctx=ctx, max_depth=8,
                mass_matrix=gm.NUTSMassMatrixConfig("dense", *cfg)).set_math_mode(exact)
    s.inject_streams(normals, exp1, unif)
    out = s.run(n_collect, n_discard)
    st = s.state()
    inv, n_updates = s.mass_matrix()
    assert inv.shape == (Cn, d, d) and n_updates == n_upd
    assert np.array_equal(st["used"][:, 0].astype(np.int64), ref["used"][:, 0])
    assert (ref["used"][:, 0] == d * (1 + n_collect + n_discard - 1 + n_upd)).all()
    same = (st["leapfrogs"] == ref["leapfrogs"])
    assert same.mean() > (0.9 if dtype == np.float64 else 0.6)
    scale = np.abs(ref["mass_inv"][same]).reshape(same.sum(), -1).max(1)
    err = np.abs(inv[same] - ref["mass_inv"][same]).reshape(same.sum(), -1).max(1) / scale
    med_tol, max_tol = (1e-8, 1e-3) if dtype == np.float64 else (2e-3, 0.5)
    print("dense mass %s %s: same trees %.3f, median inv err %.2e, max %.2e" % (dtype.__name__, "exact" if exact else "fast",
                                                                                same.mean(), np.median(err), err.max()))
    assert np.median(err) < med_tol and err.max() < max_tol
    eps_err = np.abs(st["eps"][same] - ref["eps"][same]) / ref["eps"][same]
    assert np.median(eps_err) < med_tol and eps_err.max() < max_tol
    samp_err = np.abs(out[same] - ref["samples"][same]).reshape(same.sum(), -1).max(1)
    assert np.median(samp_err) < 100 * med_tol
    assert np.allclose(inv, np.swapaxes(inv, 1, 2), rtol=1e-5, atol=1e-6)           # symmetric
    assert (np.linalg.eigvalsh(inv.astype(np.float64)) > 0).all()                   # positive definite
    assert np.abs(inv[:, 0, 1]).max() > 1e-3                                        # the off-diagonal did move


def test_nuts_dense_mass_above_dense_max_dim_keeps_identity(ctx, oracle):
    """Reference quirk, reproduced: with dim > dense_max_dim the chain keeps diagonal statistics (generic_nuts.rs:612-617)
    but maybe_update_mass_matrix still dispatches on the configured Dense adaptation, finds no dense sums and returns None
    (:972-974): the mass matrix never leaves the identity.  Bit-identical to a run without adaptation (exact mode)."""
    Cn, d, n_collect, n_discard = 64, 5, 5, 30
    tgt, cov = _corr_gauss(d)
    q0 = np.random.default_rng(2).standard_normal((Cn, d))
    normals, exp1, unif = _streams(Cn, d, n_collect + n_discard + 4, seed=29, n_unif=20000)
    ref = oracle.nuts_run(tgt.kind, tgt.params(), q0, 0.8, 8, -1.0, n_collect, n_discard, normals, exp1, unif,
                          mass_cfg=(1, 1, 10, 0.05, 1e-6, 1, 3))
    plain = oracle.nuts_run(tgt.kind, tgt.params(), q0, 0.8, 8, -1.0, n_collect, n_discard, normals, exp1, unif)
    assert (ref["mass_updates"] == 0).all() and np.array_equal(ref["samples"], plain["samples"])
    s = gm.NUTS(tgt, q0, 0.8, seed=1, ctx=ctx, max_depth=8,
                mass_matrix=gm.NUTSMassMatrixConfig("dense", 1, 1, 10, 0.05, 1e-6, dense_max_dim=3)).set_math_mode(True)
    s.inject_streams(normals, exp1, unif)
    out = s.run(n_collect, n_discard)
    p = gm.NUTS(tgt, q0, 0.8, seed=1, ctx=ctx, max_depth=8).set_math_mode(True)
    p.inject_streams(normals, exp1, unif)
    assert np.array_equal(out, p.run(n_collect, n_discard))
    assert np.array_equal(s.state()["used"].astype(np.int64), ref["used"])


def test_nuts_dense_mass_philox_recovers_correlated_gaussian(ctx, oracle):
    """Philox path, reference default windows, 300 warm-up transitions on a correlated 6-D Gaussian: the adapted matrices
    (`inv` = inverse of the regularised window covariance: the reference takes the mass to be the covariance, as its
    diagonal variant takes it to be the variance) and the draws reproduce what the reference ALGORITHM produces (the
    oracle on its own streams) — including its quirks (step-size search and sub-tree U-turn tests with the identity)."""
    Cn, d = 1024, 6
    tgt, cov = _corr_gauss(d, seed=9)
    q0 = (np.random.default_rng(5).standard_normal((Cn, d)) @ np.linalg.cholesky(cov).T).astype(np.float32)
    r2 = np.random.default_rng(1)
    Co = 384
    streams = (r2.standard_normal((Co, d * 412)), r2.exponential(size=(Co, 402)), r2.random((Co, 400000)))
    ref = oracle.nuts_run(tgt.kind, tgt.params(), q0[:Co], 0.8, 10, -1.0, 100, 300, *streams, mass_cfg=(75, 50, 25, 0.05, 1e-6, 1, 75))
    assert not ref["exhausted"].any()
    s = gm.NUTS(tgt, q0, 0.8, seed=11, ctx=ctx, max_depth=10, mass_matrix=gm.NUTSMassMatrixConfig("dense"))
    out = s.run(100, 300)
    inv, n = s.mass_matrix()
    assert n == 4 and np.isfinite(inv).all() and np.isfinite(out).all()
    got = np.median(inv.astype(np.float64), axis=0)
    want = np.median(ref["mass_inv"].astype(np.float64), axis=0)
    print("dense mass philox: median inverse mass (GPU)\n", got.round(3), "\n(oracle)\n", want.round(3), "\n(target cov)\n", cov.round(3))
    assert np.abs(got - want).max() < 0.25 * np.abs(want).max()
    cov_gpu = np.cov(out.reshape(-1, d).astype(np.float64).T)
    cov_ref = np.cov(ref["samples"].reshape(-1, d).astype(np.float64).T)
    assert np.abs(cov_gpu - cov_ref).max() < 0.2 * np.abs(cov_ref).max()
    leap_gpu = s.counters().grad_evals / (Cn * 399.0)
    leap_ref = ref["leapfrogs"].mean() / 399.0
    assert abs(leap_gpu / leap_ref - 1.0) < 0.2, (leap_gpu, leap_ref)


# -------------------------------------------------------------------------------------------------
# BASELINE config 5: NUTS on the 100-D, 4-component isotropic Gaussian mixture.  d = 100 runs at the production
# decomposition (8 coordinates x 16 lanes per chain): mixture means staged in shared memory, and in fast mode the
# packed transpose-reduce of the four components' partial sums.
# -------------------------------------------------------------------------------------------------
def _mixture(d=100, K=4):
    mu = np.stack([(k - 1.5) * (2.0 / np.sqrt(d)) * np.ones(d) for k in range(K)])
    return gm.GaussianMixture(np.full(K, 1.0 / K), mu, 1.0)


@pytest.mark.parametrize("exact", [True, False], ids=["exact", "fast"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
def test_nuts_cfg5_mixture_first_transition_matches_oracle(ctx, oracle, dtype, exact):
    """generic_nuts.rs:755-925 on the cfg5 target.  run(2, 0) = init_chain_state (find_reasonable_epsilon, powers of
    two) + ONE transition with that step size, no dual averaging in the trajectory: the only non-IEEE operations are the
    mixture's exp / log (device libm or MUFU vs glibc), so tree shapes agree except where a slice / U-turn decision sits
    on a rounding boundary, and the matching chains agree to rounding."""
    Cn, d = 96, 100
    tgt = _mixture(d)
    rng = np.random.default_rng(31)
    q0 = rng.standard_normal((Cn, d)).astype(dtype)
    normals, exp1, unif = _streams(Cn, d, 2, seed=37, n_unif=4000)
    ref = oracle.nuts_run(tgt.kind, tgt.params(), q0, 0.8, 10, -1.0, 2, 0, normals, exp1, unif)
    assert not ref["exhausted"].any()
    s = gm.NUTS(tgt, q0, 0.8, seed=1, ctx=ctx, max_depth=10).set_math_mode(exact)
    assert (s.dim, s.n_chains) == (100, Cn)
    s.inject_streams(normals, exp1, unif)
    out = s.run(2, 0)
    st = s.state()
    same = (st["leapfrogs"] == ref["leapfrogs"])
    tol = (1e-5 if not exact else 2e-6) if dtype == np.float32 else (1e-10 if not exact else 1e-12)
    err = np.abs(out[same].astype(np.float64) - ref["samples"][same].astype(np.float64)).max() if same.any() else 0.0
    print("cfg5 mixture %s %s: same trees %.3f, mean leapfrogs %.1f, max |dx| %.2e" % (
        dtype.__name__, "exact" if exact else "fast", same.mean(), ref["leapfrogs"].mean(), err))
    assert np.array_equal(out[:, 0, :], q0)
    assert same.mean() >= 0.95
    assert np.array_equal(st["eps"][same], ref["eps"][same])          # eps_bar = 1 after a transition without warm-up
    assert np.array_equal(st["used"][same].astype(np.int64), ref["used"][same])
    assert err <= tol * (1.0 + np.abs(ref["samples"]).max())
    assert ref["leapfrogs"].mean() > 6                                  # non-trivial trees


@pytest.mark.parametrize("exact", [True, False], ids=["exact", "fast"])
def test_nuts_cfg5_mixture_with_adaptation_matches_oracle(ctx, oracle, exact):
    """Warm-up + collection on the cfg5 target, f32, injected streams, as test_nuts_with_adaptation_matches_oracle."""
    Cn, d, n_collect, n_discard = 128, 100, 3, 5
    tgt = _mixture(d)
    rng = np.random.default_rng(41)
    q0 = rng.standard_normal((Cn, d)).astype(np.float32)
    normals, exp1, unif = _streams(Cn, d, n_collect + n_discard, seed=43, n_unif=20000)
    ref = oracle.nuts_run(tgt.kind, tgt.params(), q0, 0.8, 10, -1.0, n_collect, n_discard, normals, exp1, unif)
    assert not ref["exhausted"].any()
    s = gm.NUTS(tgt, q0, 0.8, seed=1, ctx=ctx, max_depth=10).set_math_mode(exact)
    s.inject_streams(normals, exp1, unif)
    out = s.run(n_collect, n_discard)
    st = s.state()
    same = (st["leapfrogs"] == ref["leapfrogs"])
    err = np.abs(out[same] - ref["samples"][same]).reshape(same.sum(), -1).max(1)
    eps_err = np.abs(st["eps"][same] - ref["eps"][same]) / ref["eps"][same]
    print("cfg5 mixture adapt %s: same trees %.3f, mean leapfrogs/transition %.1f, median |dx| %.2e, median eps err %.2e" % (
        "exact" if exact else "fast", same.mean(), ref["leapfrogs"].mean() / (n_collect + n_discard - 1), np.median(err),
        np.median(eps_err)))
    assert same.mean() > 0.8
    # f32 dual averaging goes through exp / log / pow (device libm vs glibc, 1 ulp) and feeds every rounding difference
    # back into the next trajectory: 7 transitions in, positions agree to ~1e-4 on identical trees
    assert np.median(err) < 1e-3 and np.median(eps_err) < 1e-4


def test_nuts_cfg5_mixture_distribution(ctx):
    """Philox path at the bench shape (d = 100, K = 4, depth <= 10, target accept 0.8): the chain projected on the
    direction of the component means is the 1-D mixture of N(-3,1), N(-1,1), N(1,1), N(3,1) (mean 0, variance 6), every
    orthogonal direction is N(0,1); mode occupancies from the normal CDF; split R-hat < 1.01 (north star)."""
    from math import erf, sqrt
    Cn, d = 8192, 100
    tgt = _mixture(d)
    q0 = np.random.default_rng(51).standard_normal((Cn, d)).astype(np.float32)
    s = gm.NUTS(tgt, q0, 0.8, seed=42, ctx=ctx, max_depth=10)
    out, st = s.run_progress(300, 200)
    assert np.isfinite(out).all()
    u = np.ones(d) / np.sqrt(d)
    t = out.astype(np.float64) @ u                       # [C, n]
    flat = t.ravel()
    Phi = lambda z: 0.5 * (1.0 + erf(z / sqrt(2.0)))
    modes = np.array([-3.0, -1.0, 1.0, 3.0])
    edges = [-np.inf, -2.0, 0.0, 2.0, np.inf]
    want = np.array([np.mean([Phi(edges[i + 1] - m) - Phi(edges[i] - m) for m in modes]) for i in range(4)])
    got = np.array([np.mean((flat >= edges[i]) & (flat < edges[i + 1])) for i in range(4)])
    resid = out.astype(np.float64) - t[..., None] * u
    var_orth = (resid ** 2).sum(-1).mean() / (d - 1)
    c = s.counters()
    print("cfg5 mixture distribution: occupancy", got, "want", want, "mean %.3f var %.3f var_orth %.4f rhat max %.4f "
          "ess min %.0f accept %.3f eps %.3f leapfrogs/transition %.1f" % (
              flat.mean(), flat.var(), var_orth, st.rhat_std.max, st.ess.min, c.accept_rate, c.step_size,
              c.grad_evals / c.transitions))
    assert np.allclose(got, want, atol=0.01)
    assert abs(flat.mean()) < 0.05 and abs(flat.var() / 6.0 - 1.0) < 0.03
    assert abs(var_orth - 1.0) < 0.01
    per_coord = out.reshape(-1, d).astype(np.float64).var(0)
    assert np.allclose(per_coord, 1.0 + 5.0 / d, rtol=0.02)      # (var 6 along u) / d + (1 - 1/d) orthogonal to it
    assert st.rhat_std.max < 1.01
    # divergences (energy error > 1000) only happen while the early warm-up iterations try wild step sizes
    assert 0.5 < c.accept_rate <= 1.0 and c.divergences < 1e-3 * c.grad_evals
