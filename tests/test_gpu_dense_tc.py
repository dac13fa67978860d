"""K3 (tcgen05 dense-Gaussian HMC) against the CPU oracle: per-step equivalence within the north-star
f32 tolerance (rel 1e-5 over L leapfrog steps), identical accept decisions, padding / ragged edge cases."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import general_mcmc_b200 as gm  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    return gm.default_context()


def _dense(d, seed=0):
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.standard_normal((d, d)))
    lam = np.logspace(-1, 1, d)          # eigenvalues log-spaced in [0.1, 10]
    cov = (q * lam) @ q.T
    return gm.DenseGaussian(np.zeros(d), cov=cov)


# the last case has more row tiles (313, the final one half empty) than SMs: the persistent (row-tile group, column chunk)
# schedule of the middle launches is what runs there
@pytest.mark.parametrize("d,Cn,L,eps", [(256, 128, 8, 0.05), (300, 200, 8, 0.05), (1000, 200, 32, 0.05), (64, 77, 4, 0.1),
                                        (300, 40000, 4, 0.05)])
def test_dense_tc_per_step_equivalence(ctx, oracle, d, Cn, L, eps):
    tgt = _dense(d)
    rng = np.random.default_rng(d)
    q0 = rng.standard_normal((Cn, d)).astype(np.float32)
    mom = rng.standard_normal((1, Cn, d)).astype(np.float32)
    ln_u = np.log(rng.random((1, Cn))).astype(np.float32)
    # truth: the oracle in f64 on the same f32 inputs and the same f32-rounded parameters (the f32 oracle — the reference's own
    # f32 arithmetic, sequential 1000-term f32 sums — is itself ~4e-6 away from it at d = 1000, L = 32; printed for scale)
    params32 = np.asarray(tgt.params(), np.float32).astype(np.float64)
    ref = oracle.hmc_run(tgt.kind, params32, q0.astype(np.float64), np.float64(np.float32(eps)), L,
                         mom.astype(np.float64), ln_u.astype(np.float64), want_traj=True)
    ref32 = oracle.hmc_run(tgt.kind, tgt.params(), q0, eps, L, mom, ln_u, want_traj=True)
    f32_q = np.max(np.abs(ref32["prop_q"] - ref["prop_q"]))
    f32_p = np.max(np.abs(ref32["prop_p"] - ref["prop_p"]))
    s = gm.HMC(tgt, q0, eps, L, seed=1, ctx=ctx)
    s.inject(mom, ln_u)
    out = s.run(1, 0)
    diag = s.diagnostics()
    scale_q = np.abs(ref["prop_q"]).max()
    scale_p = np.abs(ref["prop_p"]).max()
    err_q = np.max(np.abs(diag["prop_q"] - ref["prop_q"]))
    err_p = np.max(np.abs(diag["prop_p"] - ref["prop_p"]))
    print("d=%d L=%d: GPU err q %.2e p %.2e (rel %.2e %.2e); f32 reference err q %.2e p %.2e" % (
        d, L, err_q, err_p, err_q / scale_q, err_p / scale_p, f32_q, f32_p))
    # Bar: rel 1e-5 over L leapfrog steps (north star), at every shape including d = 1000, L = 32 (measured 4.6e-6 / 6.4e-6
    # since delta is scaled into FP16's normal range per transition: its low parts no longer fall into the subnormals)
    tol = 1e-5
    assert err_q <= tol * scale_q
    assert err_p <= tol * scale_p
    escale = np.abs(ref["logp_cur"][0]) + np.abs(ref["logp_prop"][0]) + 0.5 * (mom[0].astype(np.float64) ** 2).sum(-1) + 1.0
    err = np.abs(diag["log_accept"][0].astype(np.float64) - ref["log_accept"][0])
    assert np.all(err <= 4e-5 * escale), (err / escale).max()
    safe = np.abs(ref["log_accept"][0] - ln_u[0]) > 4e-5 * escale
    assert np.array_equal(diag["accepted"][0][safe], ref["accepted"][0][safe])
    assert safe.mean() > 0.97
    acc = diag["accepted"][0].astype(bool)
    assert np.array_equal(out[acc, 0], diag["prop_q"][0][acc])
    assert np.array_equal(out[~acc, 0], q0[~acc])


def test_dense_tc_bench_grid_persistent_units(ctx, oracle):
    """The bench shape's code path: d = 1000 with more row tiles (321, the last one ragged) than SMs, so the L - 1 middle
    launches run the persistent (row-tile group, column chunk) schedule.  Every chain runs on the GPU; a spread of chains
    (first / middle / last ragged tile included) is checked against the f64 oracle to the 1e-5 bar."""
    d, Cn, L, eps = 1000, 41000, 8, 0.05
    tgt = _dense(d)
    rng = np.random.default_rng(17)
    q0 = rng.standard_normal((Cn, d)).astype(np.float32)
    mom = rng.standard_normal((1, Cn, d)).astype(np.float32)
    ln_u = np.log(rng.random((1, Cn))).astype(np.float32)
    s = gm.HMC(tgt, q0, eps, L, seed=1, ctx=ctx)
    s.inject(mom, ln_u)
    out = s.run(1, 0)
    diag = s.diagnostics()
    idx = np.unique(np.concatenate([np.arange(0, Cn, 131), np.arange(Cn - 40, Cn), np.arange(20480, 20480 + 40)]))
    params32 = np.asarray(tgt.params(), np.float32).astype(np.float64)
    ref = oracle.hmc_run(tgt.kind, params32, q0[idx].astype(np.float64), np.float64(np.float32(eps)), L,
                         mom[:, idx].astype(np.float64), ln_u[:, idx].astype(np.float64), want_traj=True)
    err_q = np.abs(diag["prop_q"][0][idx] - ref["prop_q"][0]).max() / np.abs(ref["prop_q"]).max()
    err_p = np.abs(diag["prop_p"][0][idx] - ref["prop_p"][0]).max() / np.abs(ref["prop_p"]).max()
    print("bench grid d=%d C=%d L=%d (%d chains checked): rel err q %.2e p %.2e" % (d, Cn, L, idx.size, err_q, err_p))
    assert err_q <= 1e-5 and err_p <= 1e-5
    assert np.isfinite(out).all()
    escale = np.abs(ref["logp_cur"][0]) + np.abs(ref["logp_prop"][0]) + 0.5 * (mom[0][idx].astype(np.float64) ** 2).sum(-1) + 1.0
    safe = np.abs(ref["log_accept"][0] - ln_u[0][idx]) > 4e-5 * escale
    assert np.array_equal(diag["accepted"][0][idx][safe], ref["accepted"][0][safe])


@pytest.mark.parametrize("scale", [1e-8, 1.0, 1e10], ids=["cov1e-8", "cov1", "cov1e+10"])
def test_dense_tc_any_variance_scale(ctx, oracle, scale):
    """The FP16 x 3 operand split must not depend on the target's units: delta = q - mu is scaled per transition by a
    power of two (as P is at set-up), so covariance scales of 1e-8 and 1e+10 (|delta| ~ 1e-4 and ~ 1e+5, beyond FP16's
    normal range without the scale) meet the same relative bar as O(1) targets."""
    d, Cn, L = 128, 256, 8
    rng = np.random.default_rng(5)
    qm, _ = np.linalg.qr(rng.standard_normal((d, d)))
    lam = np.logspace(-1, 1, d) * scale
    tgt = gm.DenseGaussian(rng.standard_normal(d) * np.sqrt(scale), cov=(qm * lam) @ qm.T)
    eps = 0.05 * np.sqrt(scale)
    q0 = (tgt.mean + rng.standard_normal((Cn, d)) * np.sqrt(scale)).astype(np.float32)
    # with eps ~ sqrt(scale) the dynamics are scale-free: momenta stay N(0, 1) (kinetic and potential energy are O(1))
    mom = rng.standard_normal((1, Cn, d)).astype(np.float32)
    ln_u = np.log(rng.random((1, Cn))).astype(np.float32)
    params32 = np.asarray(tgt.params(), np.float32).astype(np.float64)
    ref = oracle.hmc_run(tgt.kind, params32, q0.astype(np.float64), np.float64(np.float32(eps)), L,
                         mom.astype(np.float64), ln_u.astype(np.float64), want_traj=True)
    s = gm.HMC(tgt, q0, eps, L, seed=1, ctx=ctx)
    s.inject(mom, ln_u)
    s.run(1, 0)
    diag = s.diagnostics()
    assert np.isfinite(diag["prop_q"]).all() and np.isfinite(diag["log_accept"]).all()
    dq = ref["prop_q"] - q0.astype(np.float64)[None]
    err_q = np.max(np.abs(diag["prop_q"] - ref["prop_q"])) / max(np.abs(dq).max(), np.abs(q0 - tgt.mean).max())
    err_p = np.max(np.abs(diag["prop_p"] - ref["prop_p"])) / np.abs(ref["prop_p"]).max()
    print("cov scale %g: rel err q %.2e p %.2e, accept %.2f" % (scale, err_q, err_p, diag["accepted"].mean()))
    assert err_q <= 1e-5 and err_p <= 1e-5
    assert diag["accepted"].mean() > 0.3


def test_dense_tc_matches_register_kernel_and_distribution(ctx):
    """Same Philox streams as K1: the tensor-core path and the register path (GMCMC_DENSE_TC=0 is the env switch;
    here the exact-mode sampler uses K1) produce the same chains up to the 3xTF32 rounding, and recover the target
    covariance."""
    d, Cn = 64, 2048
    tgt = _dense(d, seed=3)
    q0 = np.zeros((Cn, d), np.float32)
    s = gm.HMC(tgt, q0, 0.15, 8, seed=42, ctx=ctx)
    out = s.run(200, 100)
    flat = out[:, ::10].reshape(-1, d).astype(np.float64)
    cov = np.cov(flat.T)
    true = np.linalg.inv(tgt.precision)
    assert np.allclose(flat.mean(0), 0.0, atol=0.15)
    assert np.abs(cov - true).max() < 0.12 * np.abs(true).max()
    c = s.counters()
    assert 0.6 < c.accept_rate <= 1.0
