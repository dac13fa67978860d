"""ctypes binding of the CPU oracle (oracle/liboracle.so).  Test infrastructure only.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module; the
product package (general_mcmc_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")

# target kinds (mirror include/gmcmc.h)
ISO_GAUSS, GAUSS2D, DIFF_GAUSS2D, DENSE_GAUSS, ROSENBROCK2D, ROSENBROCK_ND, GAUSS_MIXTURE = range(7)

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C")


def build(force=False):
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    fast = os.path.join(ORACLE_DIR, "liboracle_fast.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("oracle_capi.cpp", "gmcmc_oracle.hpp")]
    stale = force or not (os.path.exists(so) and os.path.exists(fast))
    if not stale:
        t = min(os.path.getmtime(so), os.path.getmtime(fast))
        stale = any(os.path.getmtime(s) > t for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return so, fast


_libs = {}


def lib(fast=False):
    key = "fast" if fast else "exact"
    if key not in _libs:
        so, fso = build()
        L = C.CDLL(fso if fast else so)
        L.orc_target_logp_grad_f64.restype = C.c_double
        L.orc_target_logp_grad_f32.restype = C.c_float
        L.orc_target_logp_f64.restype = C.c_double
        L.orc_target_logp_f32.restype = C.c_float
        L.orc_iso_proposal_logp_f64.restype = C.c_double
        L.orc_nuts_find_reasonable_epsilon_f64.restype = C.c_double
        L.orc_nuts_find_reasonable_epsilon_f32.restype = C.c_float
        L.orc_hmc_bench_f32.restype = C.c_double
        L.orc_hmc_bench_f64.restype = C.c_double
        L.orc_mh_bench_f64.restype = C.c_double
        L.orc_max_threads.restype = C.c_int
        L.orc_diag_mass_kinetic_inv_mul_f64.restype = C.c_double
        L.orc_int_target_logp.restype = C.c_double
        _libs[key] = L
    return _libs[key]


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _params(params):
    p = np.ascontiguousarray(np.asarray(params, dtype=np.float64).ravel())
    return p, _p(p), C.c_size_t(p.size)


def _sfx(dtype):
    return "f32" if np.dtype(dtype) == np.float32 else "f64"


def _ct(dtype):
    return C.c_float if np.dtype(dtype) == np.float32 else C.c_double


def target_logp_grad(kind, dim, params, x):
    x = np.ascontiguousarray(x)
    g = np.zeros_like(x)
    p, pp, n = _params(params)
    f = getattr(lib(), "orc_target_logp_grad_" + _sfx(x.dtype))
    lp = f(C.c_int(kind), C.c_int(dim), pp, n, _p(x), _p(g))
    return lp, g


def target_logp(kind, dim, params, x):
    x = np.ascontiguousarray(x)
    p, pp, n = _params(params)
    f = getattr(lib(), "orc_target_logp_" + _sfx(x.dtype))
    return f(C.c_int(kind), C.c_int(dim), pp, n, _p(x))


def iso_proposal_logp(frm, to, std):
    frm = np.ascontiguousarray(frm, np.float64)
    to = np.ascontiguousarray(to, np.float64)
    return lib().orc_iso_proposal_logp_f64(_p(frm), _p(to), C.c_int(frm.size), C.c_double(std))


def hmc_run(kind, params, q0, eps, L, momenta, ln_u, want_traj=False):
    """q0 [C,d]; momenta [n,C,d]; ln_u [n,C].  Returns dict with final q, samples [C,n,d], accepted
    [n,C], log_accept [n,C] (+ prop_q/prop_p [n,C,d], logp_cur/logp_prop when want_traj)."""
    q = np.array(q0, copy=True, order="C")
    dt = q.dtype
    Cn, d = q.shape
    momenta = np.ascontiguousarray(momenta, dt)
    ln_u = np.ascontiguousarray(ln_u, dt)
    n = momenta.shape[0]
    samples = np.zeros((Cn, n, d), dt)
    acc = np.zeros((n, Cn), np.uint8)
    la = np.zeros((n, Cn), dt)
    pq = np.zeros((n, Cn, d), dt) if want_traj else None
    pp_ = np.zeros((n, Cn, d), dt) if want_traj else None
    lc = np.zeros((n, Cn), dt) if want_traj else None
    lpr = np.zeros((n, Cn), dt) if want_traj else None
    p, ppar, np_ = _params(params)
    f = getattr(lib(), "orc_hmc_run_" + _sfx(dt))
    f(C.c_int(kind), C.c_int(d), ppar, np_, C.c_size_t(Cn), _p(q), _ct(dt)(eps), C.c_int(L), C.c_size_t(n),
      _p(momenta), _p(ln_u), _p(samples), _p(acc), _p(la), _p(pq), _p(pp_), _p(lc), _p(lpr))
    return dict(q=q, samples=samples, accepted=acc, log_accept=la, prop_q=pq, prop_p=pp_, logp_cur=lc,
                logp_prop=lpr)


def mh_run(kind, params, x0, prop_std, normals, ln_u):
    x = np.array(x0, copy=True, order="C")
    dt = x.dtype
    Cn, d = x.shape
    normals = np.ascontiguousarray(normals, dt)
    ln_u = np.ascontiguousarray(ln_u, dt)
    n = normals.shape[0]
    samples = np.zeros((Cn, n, d), np.float64)
    acc = np.zeros((n, Cn), np.uint8)
    lr = np.zeros((n, Cn), dt)
    p, ppar, np_ = _params(params)
    f = getattr(lib(), "orc_mh_run_" + _sfx(dt))
    f(C.c_int(kind), C.c_int(d), ppar, np_, _ct(dt)(prop_std), C.c_size_t(Cn), _p(x), C.c_size_t(n), _p(normals),
      _p(ln_u), _p(samples), _p(acc), _p(lr))
    return dict(x=x, samples=samples, accepted=acc, log_ratio=lr)


def nuts_build_tree(kind, params, q, p, g, logu, v, j, eps, joint_0, unif=()):
    q = np.ascontiguousarray(q)
    dt = q.dtype
    d = q.size
    p = np.ascontiguousarray(p, dt)
    g = np.ascontiguousarray(g, dt)
    unif = np.ascontiguousarray(np.asarray(unif, np.float64))
    vecs = np.zeros((8, d), dt)
    sc = np.zeros(2, dt)
    ints = np.zeros(5, np.int64)
    par, ppar, np_ = _params(params)
    f = getattr(lib(), "orc_nuts_build_tree_" + _sfx(dt))
    ct = _ct(dt)
    f(C.c_int(kind), C.c_int(d), ppar, np_, _p(q), _p(p), _p(g), ct(logu), C.c_int(v), C.c_int(j), ct(eps),
      ct(joint_0), _p(unif), C.c_size_t(unif.size), _p(vecs), _p(sc), _p(ints))
    names = ["q_minus", "p_minus", "g_minus", "q_plus", "p_plus", "g_plus", "q_prime", "g_prime"]
    out = {k: vecs[i] for i, k in enumerate(names)}
    out.update(logp_prime=sc[0], alpha_prime=sc[1], n_prime=int(ints[0]), s_prime=bool(ints[1]),
               n_alpha_prime=int(ints[2]), leapfrogs=int(ints[3]), unif_used=int(ints[4]))
    return out


def nuts_find_reasonable_epsilon(kind, params, q, p):
    q = np.ascontiguousarray(q)
    p = np.ascontiguousarray(p, q.dtype)
    par, ppar, np_ = _params(params)
    f = getattr(lib(), "orc_nuts_find_reasonable_epsilon_" + _sfx(q.dtype))
    return f(C.c_int(kind), C.c_int(q.size), ppar, np_, _p(q), _p(p))


def nuts_run(kind, params, q0, target_accept, max_depth, eps_init, n_collect, n_discard, normals, exp1, unif,
             mass_cfg=None, fast=False):
    """normals [C,nn], exp1 [C,ne], unif [C,nu] float64 streams (consumed in reference order).  mass_cfg =
    (start_buffer, end_buffer, initial_window, regularize, jitter) enables diagonal mass-matrix adaptation."""
    q = np.array(q0, copy=True, order="C")
    dt = q.dtype
    Cn, d = q.shape
    normals = np.ascontiguousarray(normals, np.float64)
    exp1 = np.ascontiguousarray(exp1, np.float64)
    unif = np.ascontiguousarray(unif, np.float64)
    samples = np.zeros((Cn, n_collect, d), dt)
    eps_f = np.zeros(Cn, dt)
    leap = np.zeros(Cn, np.int64)
    used = np.zeros((Cn, 3), np.int64)
    exh = np.zeros(Cn, np.int32)
    par, ppar, np_ = _params(params)
    ct = _ct(dt)
    args = [C.c_int(kind), C.c_int(d), ppar, np_, C.c_size_t(Cn), _p(q), ct(target_accept), C.c_int(max_depth),
            ct(eps_init), C.c_size_t(n_collect), C.c_size_t(n_discard), _p(normals), C.c_size_t(normals.shape[1]),
            _p(exp1), C.c_size_t(exp1.shape[1]), _p(unif), C.c_size_t(unif.shape[1]), _p(samples), _p(eps_f),
            _p(leap), _p(used), _p(exh)]
    mass_inv = None
    mass_updates = None
    if mass_cfg is None:
        getattr(lib(fast), "orc_nuts_run_" + _sfx(dt))(*args)
    else:
        # (start_buffer, end_buffer, initial_window, regularize, jitter [, dense (0 / 1), dense_max_dim])
        cfg = np.zeros(7, np.float64)
        cfg[6] = 75
        cfg[:len(mass_cfg)] = mass_cfg
        dense = cfg[5] != 0.0
        mass_inv = np.ones((Cn, d, d) if dense else (Cn, d), dt)
        mass_updates = np.zeros(Cn, np.int64)
        getattr(lib(fast), "orc_nuts_run_mass_" + _sfx(dt))(*args, _p(cfg), _p(mass_inv), _p(mass_updates))
    return dict(q=q, samples=samples, eps=eps_f, leapfrogs=leap, used=used, exhausted=exh, mass_inv=mass_inv,
                mass_updates=mass_updates)


POISSON, BINOMIAL = 0, 1


def mh_int_run(kind, params, x0, steps, ln_u):
    """Integer-state MH (tests/metrohast_poisson_test.rs): x0 [C,d] int32, steps [n,C,d] int8 (+-1), ln_u [n,C] f64."""
    x = np.array(x0, dtype=np.int32, copy=True, order="C")
    Cn, d = x.shape
    steps = np.ascontiguousarray(steps, np.int8)
    ln_u = np.ascontiguousarray(ln_u, np.float64)
    n = steps.shape[0]
    samples = np.zeros((Cn, n, d), np.float64)
    acc = np.zeros((n, Cn), np.uint8)
    lr = np.zeros((n, Cn), np.float64)
    p = np.ascontiguousarray(params, np.float64)
    lib().orc_mh_int_run(C.c_int(kind), C.c_int(d), _p(p), C.c_size_t(Cn), _p(x), C.c_size_t(n), _p(steps), _p(ln_u),
                         _p(samples), _p(acc), _p(lr))
    return dict(x=x, samples=samples, accepted=acc, log_ratio=lr)


def gibbs_run(kind, params, x0, normals, uniforms):
    """Gibbs sweeps (gibbs.rs:89-105) with the reference tests' conditionals: x0 [C,d] f64, normals / uniforms [n,C,d]."""
    x = np.array(x0, dtype=np.float64, copy=True, order="C")
    Cn, d = x.shape
    normals = np.ascontiguousarray(normals, np.float64)
    uniforms = np.ascontiguousarray(uniforms, np.float64)
    n = normals.shape[0]
    samples = np.zeros((Cn, n, d), np.float64)
    p = np.ascontiguousarray(params, np.float64)
    lib().orc_gibbs_run(C.c_int(kind), C.c_int(d), _p(p), C.c_size_t(Cn), _p(x), C.c_size_t(n), _p(normals), _p(uniforms), _p(samples))
    return dict(x=x, samples=samples)


def int_target_logp(kind, params, k):
    k = np.ascontiguousarray(k, np.int32)
    p = np.ascontiguousarray(params, np.float64)
    return lib().orc_int_target_logp(C.c_int(kind), C.c_int(k.size), _p(p), _p(k))


def dense_mass(cov, p, jitter=1e-12):
    """MassMatrix::dense_from_cov(cov, d, jitter) -> dict(inv_mul(p), inv, chol, kinetic(p)) or None (generic_nuts.rs:208-359)."""
    cov = np.ascontiguousarray(cov, np.float64)
    d = cov.shape[0]
    p = np.ascontiguousarray(p, np.float64)
    out, inv, chol = np.zeros(d), np.zeros((d, d)), np.zeros((d, d))
    ke = C.c_double(0)
    ok = lib().orc_dense_mass_inv_mul_f64(_p(cov), C.c_int(d), C.c_double(jitter), _p(p), _p(out), _p(inv), _p(chol), C.byref(ke))
    return dict(inv_mul=out, inv=inv, chol=chol, kinetic=ke.value) if ok else None


def diag_mass_kinetic_inv_mul(var, p, jitter=1e-12):
    """MassMatrix::diagonal_from_var(var, jitter) -> (kinetic(p), inv_mul(p)) (generic_nuts.rs:196-281)."""
    var = np.ascontiguousarray(var, np.float64)
    p = np.ascontiguousarray(p, np.float64)
    out = np.zeros_like(p)
    ke = lib().orc_diag_mass_kinetic_inv_mul_f64(_p(var), C.c_int(var.size), C.c_double(jitter), _p(p), _p(out))
    return ke, out


def split_rhat_mean_ess(sample):
    s = np.ascontiguousarray(sample, np.float32)
    c, n, p = s.shape
    rhat = np.zeros(p, np.float32)
    ess = np.zeros(p, np.float32)
    lib().orc_split_rhat_mean_ess(_p(s), C.c_size_t(c), C.c_size_t(n), C.c_size_t(p), _p(rhat), _p(ess))
    return rhat, ess


def autocov(x, fft):
    x = np.ascontiguousarray(x, np.float32)
    n, d = x.shape
    out = np.zeros((n, d), np.float32)
    f = lib().orc_autocov_fft if fft else lib().orc_autocov_bf
    f(_p(x), C.c_size_t(n), C.c_size_t(d), _p(out))
    return out


def basic_stats(data):
    data = np.ascontiguousarray(data, np.float32)
    out = np.zeros(5, np.float32)
    lib().orc_basic_stats(_p(data), C.c_size_t(data.size), _p(out))
    return dict(min=out[0], median=out[1], max=out[2], mean=out[3], std=out[4])


def tracker_rhat(states):
    s = np.ascontiguousarray(states, np.float32)
    steps, c, p = s.shape
    rhat = np.zeros(p, np.float32)
    pa = C.c_float(0)
    lib().orc_tracker_rhat(_p(s), C.c_size_t(steps), C.c_size_t(c), C.c_size_t(p), _p(rhat), C.byref(pa))
    return rhat, pa.value


def philox4x32_10(ctr, key):
    ctr = np.ascontiguousarray(ctr, np.uint32)
    key = np.ascontiguousarray(key, np.uint32)
    out = np.zeros(4, np.uint32)
    lib().orc_philox4x32_10(_p(ctr), _p(key), _p(out))
    return out


def hmc_bench(kind, params, q0, eps, L, n_steps, seed=42, threads=0, keep_samples=False):
    q = np.array(q0, copy=True, order="C")
    dt = q.dtype
    Cn, d = q.shape
    samples = np.zeros((Cn, n_steps, d), dt) if keep_samples else None
    par, ppar, np_ = _params(params)
    f = getattr(lib(fast=True), "orc_hmc_bench_" + _sfx(dt))
    secs = f(C.c_int(kind), C.c_int(d), ppar, np_, C.c_size_t(Cn), _p(q), _ct(dt)(eps), C.c_int(L),
             C.c_size_t(n_steps), C.c_uint64(seed), C.c_int(threads), _p(samples))
    return secs, q, samples


def mh_bench(kind, params, x0, prop_std, n_steps, seed=42, threads=0, keep_samples=False):
    x = np.array(x0, dtype=np.float64, copy=True, order="C")
    Cn, d = x.shape
    samples = np.zeros((Cn, n_steps, d), np.float64) if keep_samples else None
    par, ppar, np_ = _params(params)
    secs = lib(fast=True).orc_mh_bench_f64(C.c_int(kind), C.c_int(d), ppar, np_, C.c_double(prop_std),
                                           C.c_size_t(Cn), _p(x), C.c_size_t(n_steps), C.c_uint64(seed),
                                           C.c_int(threads), _p(samples))
    return secs, x, samples


def set_threads(n, fast=True):
    lib(fast).orc_set_threads(C.c_int(n))


def max_threads():
    return lib(fast=True).orc_max_threads()
