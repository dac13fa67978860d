"""Host-side mirror of the reference's sampler front-ends over the C ABI (include/gmcmc.h).

The names, argument meaning and output layout follow the Rust crate so that callers (and the parity
tests) read like the reference's own code:

  HMC(target, initial_positions, step_size, n_leapfrog)        hmc.rs:113-134
      .set_seed(seed) .run(n_collect, n_discard) .run_progress(...) .step() .positions()   hmc.rs:143-338
  MetropolisHastings(target, proposal, initial_states)         metropolis_hastings.rs:151-187
      .seed(seed) .run(...) .run_progress(...)                 metropolis_hastings.rs:189-197, core.rs:219-255
  NUTS(target, initial_positions, target_accept_p)             nuts.rs:156-304
  RunStats / BasicStats / split_rhat_mean_ess                  stats.rs:342-450
  init / init_det / init_with_seed                             core.rs:434-475

Samples come back as numpy arrays [chains, samples, dim] (core.rs:219-229, hmc.rs:179-180): MH always
float64, HMC/NUTS in the sampler's dtype.  All arithmetic happens in libgmcmc.so on the GPU.
"""
import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import _lib as L


# ------------------------------------------------------------------------------------------------
# context
# ------------------------------------------------------------------------------------------------
class Context:
    """One GPU (= one process rank).  rank/world describe the chain sharding over GPUs."""

    def __init__(self, device=0, rank=0, world=1, nccl_id=None):
        self._h = C.c_void_p()
        self.device, self.rank, self.world = device, rank, world
        if world > 1:
            if nccl_id is None or len(nccl_id) != 128:
                raise ValueError("a 128-byte NCCL unique id is required when world > 1")
            buf = (C.c_char * 128).from_buffer_copy(bytes(nccl_id))
            L.check(L.lib().gmcmc_ctx_create_dist(device, rank, world, buf, C.byref(self._h)))
        else:
            L.check(L.lib().gmcmc_ctx_create(device, C.byref(self._h)))

    @staticmethod
    def nccl_unique_id():
        buf = (C.c_char * 128)()
        L.check(L.lib().gmcmc_nccl_unique_id(buf))
        return bytes(buf)

    def synchronize(self):
        L.check(L.lib().gmcmc_ctx_synchronize(self._h))

    def stream(self):
        s = C.c_void_p()
        L.check(L.lib().gmcmc_ctx_stream(self._h, C.byref(s)))
        return s.value

    def all_reduce(self, values):
        a = np.ascontiguousarray(values, dtype=np.float64).copy()
        L.check(L.lib().gmcmc_ctx_all_reduce_f64(self._h, L.ptr(a), C.c_size_t(a.size)))
        return a

    def close(self):
        if self._h:
            L.lib().gmcmc_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def set_default_context(ctx):
    global _default_ctx
    _default_ctx = ctx


def shard_chains(n_chains, rank, world):
    """Contiguous chain range [lo, hi) owned by `rank`: Philox streams are keyed by the
    global chain index, so results do not depend on `world`."""
    lo = (n_chains * rank) // world
    hi = (n_chains * (rank + 1)) // world
    return lo, hi


# ------------------------------------------------------------------------------------------------
# targets and proposals (distributions.rs)
# ------------------------------------------------------------------------------------------------
class _Target:
    kind = None
    dim = None

    def params(self):
        raise NotImplementedError

    def _create(self, ctx, dtype):
        p = np.ascontiguousarray(self.params(), dtype=np.float64)
        h = C.c_void_p()
        L.check(L.lib().gmcmc_target_create(ctx._h, self.kind, L.dtype_code(dtype), int(self.dim), L.ptr(p),
                                            C.c_size_t(p.size), C.byref(h)))
        return h

    def logp_and_grad(self, x, ctx=None, exact=False):
        """≙ BatchedHamiltonianTarget::logp_and_grad (batched_hmc.rs:18-22) on host arrays [n, dim]."""
        ctx = ctx or default_context()
        x = np.ascontiguousarray(x)
        if self.dim is None:
            self.dim = int(x.shape[1])
        h = self._create(ctx, x.dtype)
        try:
            lp = np.empty(x.shape[0], x.dtype)
            g = np.empty_like(x)
            L.check(L.lib().gmcmc_target_logp_grad(h, L.ptr(x), C.c_size_t(x.shape[0]), L.ptr(lp), L.ptr(g),
                                                   L.MATH_EXACT if exact else L.MATH_FAST))
        finally:
            L.lib().gmcmc_target_destroy(h)
        return lp, g


class IsotropicGaussian(_Target):
    """distributions.rs:349-406 — both an MH proposal (`std` = proposal scale) and a target."""
    kind = L.TARGET_ISO_GAUSS

    def __init__(self, std, dim=None):
        self.std = float(std)
        self.dim = dim

    def params(self):
        return [self.std]


class Gaussian2D(_Target):
    """distributions.rs:161-208."""
    kind = L.TARGET_GAUSS2D
    dim = 2

    def __init__(self, mean, cov):
        self.mean = np.asarray(mean, np.float64).reshape(2)
        self.cov = np.asarray(cov, np.float64).reshape(2, 2)

    def params(self):
        return np.concatenate([self.mean, self.cov.ravel()])


class DiffableGaussian2D(Gaussian2D):
    """distributions.rs:215-320."""
    kind = L.TARGET_DIFF_GAUSS2D


class DenseGaussian(_Target):
    """N-D dense-covariance Gaussian: the N-D form of DiffableGaussian2D (distributions.rs:229-291)."""
    kind = L.TARGET_DENSE_GAUSS

    def __init__(self, mean, cov=None, precision=None):
        self.mean = np.asarray(mean, np.float64).ravel()
        self.dim = self.mean.size
        if precision is None:
            cov = np.asarray(cov, np.float64)
            precision = np.linalg.inv(cov)
            sign, logdet = np.linalg.slogdet(cov)
        else:
            precision = np.asarray(precision, np.float64)
            sign, logdet = np.linalg.slogdet(precision)
            logdet = -logdet
        self.precision = 0.5 * (precision + precision.T)
        self.norm_const = -(self.dim * np.log(2.0 * np.pi) + logdet) / 2.0

    def params(self):
        return np.concatenate([self.mean, self.precision.ravel(), [self.norm_const]])


class Rosenbrock2D(_Target):
    """distributions.rs:495-530."""
    kind = L.TARGET_ROSENBROCK2D
    dim = 2

    def __init__(self, a=1.0, b=100.0):
        self.a, self.b = float(a), float(b)

    def params(self):
        return [self.a, self.b]


class RosenbrockND(_Target):
    """distributions.rs:535-555."""
    kind = L.TARGET_ROSENBROCK_ND

    def __init__(self, dim=None):
        self.dim = dim

    def params(self):
        return []


class GaussianMixture(_Target):
    """Isotropic Gaussian mixture (synthetic target of BASELINE config 5)."""
    kind = L.TARGET_GAUSS_MIXTURE

    def __init__(self, weights, means, sigma=1.0):
        self.weights = np.asarray(weights, np.float64).ravel()
        self.means = np.asarray(means, np.float64).reshape(self.weights.size, -1)
        self.dim = self.means.shape[1]
        self.sigma = float(sigma)

    def params(self):
        return np.concatenate([[self.weights.size, self.sigma], self.weights, self.means.ravel()])


class PoissonTarget:
    """Discrete target of tests/metrohast_poisson_test.rs:18-50: unnorm_logp(k) = k ln(lambda) - lambda - ln(k!)."""
    int_kind = 0

    def __init__(self, lam):
        self.lam = float(lam)

    def params(self):
        return [self.lam]


class BinomialTarget:
    """Discrete target of tests/metrohast_poisson_test.rs:195-222: Binomial(n, p) on {0, .., n}."""
    int_kind = 1

    def __init__(self, n, p):
        self.n, self.p = int(n), float(p)

    def params(self):
        return [float(self.n), self.p]


class RandomWalkProposal:
    """The +-1 random walk of the reference's discrete tests (PoissonRandomWalk / BinomialRandomWalk,
    tests/metrohast_poisson_test.rs:52-90, 224-252): each coordinate moves up or down with probability 1/2, clamped to
    the target's support; logp(from, to) = ln 0.5."""


class CustomTarget(_Target):
    """A user-written device target compiled ahead of time into a plugin (csrc/gmcmc_custom_target.cuh);
    ≙ implementing GradientTarget / BatchedGradientTarget (distributions.rs:67-90) in the reference."""
    kind = 7

    def __init__(self, plugin_path, dim, params=()):
        self.plugin_path = os.path.abspath(plugin_path)
        self.dim = int(dim)
        self._params = np.asarray(params, np.float64).ravel()

    def params(self):
        return self._params

    def _create(self, ctx, dtype):
        p = np.ascontiguousarray(self._params, dtype=np.float64)
        h = C.c_void_p()
        L.check(L.lib().gmcmc_target_create_custom(ctx._h, self.plugin_path.encode(), L.dtype_code(dtype), L.ptr(p),
                                                   C.c_size_t(p.size), C.byref(h)))
        return h


def build_custom_target(source, out=None, extra_flags=()):
    """nvcc-compiles a custom-target source (see csrc/gmcmc_custom_target.cuh) into a plugin .so for sm_100a."""
    import shutil
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    out = out or os.path.splitext(source)[0] + ".so"
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-Xcompiler", "-fPIC",
           "-shared", "-I", os.path.join(here, "csrc"), "-I", os.path.join(os.path.dirname(here), "include"),
           *extra_flags, source, "-o", out]
    subprocess.check_call(cmd)
    return out


def build_custom_conditional(source, out=None, extra_flags=()):
    """nvcc-compiles a Gibbs conditional source (see csrc/gmcmc_custom_conditional.cuh) into a plugin .so for sm_100a."""
    return build_custom_target(source, out, ("--fmad=false",) + tuple(extra_flags))


# ------------------------------------------------------------------------------------------------
# statistics (stats.rs)
# ------------------------------------------------------------------------------------------------
@dataclass
class BasicStats:
    name: str
    min: float
    median: float
    max: float
    mean: float
    std: float

    def __str__(self):  # stats.rs:406-415
        return "%s in [%.2f, %.2f], median: %.2f, mean: %.2f ± %.2f" % (
            self.name, self.min, self.max, self.median, self.mean, self.std)


@dataclass
class RunStats:
    ess: BasicStats
    rhat: BasicStats            # reference orientation sqrt(W / var_hat), stats.rs:452-454
    rhat_std: BasicStats = None  # standard orientation sqrt(var_hat / W)

    def __str__(self):
        return "%s\n%s" % (self.ess, self.rhat)

    @staticmethod
    def _from_c(c):
        def b(name, s):
            return BasicStats(name, s.min, s.median, s.max, s.mean, s.std)
        return RunStats(b("ESS", c.ess), b("Split R-hat", c.rhat), b("Split R-hat (var/W)", c.rhat_std))

    @staticmethod
    def from_samples(sample, ctx=None):
        """≙ RunStats::from (stats.rs:383-394), computed on the device."""
        ctx = ctx or default_context()
        s = np.ascontiguousarray(sample)
        if s.dtype not in (np.float32, np.float64):
            s = s.astype(np.float64)
        c, n, p = s.shape
        out = L.RunStatsC()
        L.check(L.lib().gmcmc_run_stats_from(ctx._h, L.ptr(s), C.c_size_t(c), C.c_size_t(n), C.c_size_t(p),
                                             L.dtype_code(s.dtype), 0, C.byref(out)))
        return RunStats._from_c(out)


def split_rhat_mean_ess(sample, ctx=None):
    """≙ stats::split_rhat_mean_ess (stats.rs:439-450): returns (rhat[p], ess[p]) as float32."""
    ctx = ctx or default_context()
    s = np.ascontiguousarray(sample)
    if s.dtype not in (np.float32, np.float64):
        s = s.astype(np.float64)
    c, n, p = s.shape
    rhat = np.empty(p, np.float32)
    ess = np.empty(p, np.float32)
    L.check(L.lib().gmcmc_split_rhat_ess(ctx._h, L.ptr(s), C.c_size_t(c), C.c_size_t(n), C.c_size_t(p),
                                         L.dtype_code(s.dtype), 0, L.ptr(rhat), L.ptr(ess)))
    return rhat, ess


def tracker_stats(sample, ctx=None):
    """≙ MultiChainTracker (stats.rs:199-339) stepped once per draw of `sample` [chains, draws, params]: returns
    {"rhat": float32[p], "max_rhat": float, "p_accept": float} — the figures run_progress displays."""
    ctx = ctx or default_context()
    s = np.ascontiguousarray(sample)
    if s.dtype not in (np.float32, np.float64):
        s = s.astype(np.float64)
    c, n, p = s.shape
    rhat = np.empty(p, np.float32)
    mx = C.c_float(0)
    pa = C.c_float(0)
    L.check(L.lib().gmcmc_tracker_stats(ctx._h, L.ptr(s), C.c_size_t(c), C.c_size_t(n), C.c_size_t(p),
                                        L.dtype_code(s.dtype), 0, L.ptr(rhat), C.byref(mx), C.byref(pa)))
    return {"rhat": rhat, "max_rhat": float(mx.value), "p_accept": float(pa.value)}


@dataclass
class Counters:
    transitions: int
    accepts: int
    grad_evals: int
    divergences: int
    step_size: float
    kernel_ms: float
    launches: int

    @property
    def accept_rate(self):
        return self.accepts / max(1, self.transitions)


# ------------------------------------------------------------------------------------------------
# samplers
# ------------------------------------------------------------------------------------------------
def _as_positions(initial, dtype=None):
    a = np.asarray(initial)
    if a.ndim != 2:
        raise ValueError("initial positions must be [n_chains, dim]")
    if dtype is None:
        dtype = a.dtype if a.dtype in (np.float32, np.float64) else np.float64
    return np.ascontiguousarray(a, dtype=dtype)


class _Sampler:
    _h = None
    _out_dtype = None
    _n_inj = 0

    def _info(self):
        return self.n_chains, self.dim

    def _destroy(self):
        if self._h:
            L.lib().gmcmc_sampler_destroy(self._h)
            self._h = None
        if getattr(self, "_th", None):
            L.lib().gmcmc_target_destroy(self._th)
            self._th = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    # -- reference surface ---------------------------------------------------------------------
    def run(self, n_collect, n_discard=0, out=None):
        """[n_chains, n_collect, dim] samples after n_discard burn-in transitions."""
        if out is None:
            out = np.empty((self.n_chains, n_collect, self.dim), self._out_dtype)
        assert out.shape == (self.n_chains, n_collect, self.dim) and out.flags.c_contiguous
        L.check(L.lib().gmcmc_run(self._h, C.c_size_t(n_collect), C.c_size_t(n_discard), L.ptr(out),
                                  L.dtype_code(out.dtype)))
        return out

    def run_progress(self, n_collect, n_discard=0, want_samples=True):
        """(samples, RunStats) like the reference's run_progress, without the terminal UI; the
        statistics are reduced on the device over ALL ranks' chains."""
        out = np.empty((self.n_chains, n_collect, self.dim), self._out_dtype) if want_samples else None
        st = L.RunStatsC()
        L.check(L.lib().gmcmc_run_stats(self._h, C.c_size_t(n_collect), C.c_size_t(n_discard), L.ptr(out),
                                        L.dtype_code(self._out_dtype), C.byref(st)))
        return out, RunStats._from_c(st)

    def run_device(self, n_collect, n_discard=0):
        """≙ run_positions (batched_hmc.rs:115-123): samples stay on the GPU; returns the device
        pointer (int) of the library-owned [n_chains, n_collect, dim] tensor."""
        p = C.c_void_p()
        L.check(L.lib().gmcmc_run_device(self._h, C.c_size_t(n_collect), C.c_size_t(n_discard), C.byref(p)))
        return p.value

    def reserve(self, n_collect):
        """Sizes the device sample buffer for up to n_collect draws per chain ahead of the run (0 frees it)."""
        L.check(L.lib().gmcmc_reserve_samples(self._h, C.c_size_t(n_collect)))

    def close(self):
        """Frees the sampler's device memory now (≙ drop)."""
        self._destroy()

    def step(self):
        L.check(L.lib().gmcmc_step(self._h))

    def positions(self):
        out = np.empty((self.n_chains, self.dim), self.dtype)
        L.check(L.lib().gmcmc_positions(self._h, L.ptr(out)))
        return out

    def set_positions(self, positions):
        a = (np.ascontiguousarray(positions, np.int32) if getattr(self, "_discrete", False)
             else _as_positions(positions, self.dtype))
        assert a.shape == (self.n_chains, self.dim)
        L.check(L.lib().gmcmc_set_positions(self._h, L.ptr(a)))
        self.ctx.synchronize()

    def set_seed(self, seed):
        L.check(L.lib().gmcmc_set_seed(self._h, C.c_uint64(seed)))
        return self

    # -- extensions -----------------------------------------------------------------------------
    def set_math_mode(self, exact):
        L.check(L.lib().gmcmc_set_math_mode(self._h, L.MATH_EXACT if exact else L.MATH_FAST))
        return self

    def inject(self, normals, ln_u):
        """Test hook: the next len(ln_u) transitions consume these draws instead of Philox."""
        normals = np.ascontiguousarray(normals, self.dtype)
        ln_u = np.ascontiguousarray(ln_u, self.dtype)
        n = ln_u.shape[0]
        assert normals.shape == (n, self.n_chains, self.dim) and ln_u.shape == (n, self.n_chains)
        L.check(L.lib().gmcmc_inject(self._h, L.ptr(normals), L.ptr(ln_u), C.c_size_t(n)))
        self._n_inj = n

    def diagnostics(self):
        n = self._n_inj
        if n == 0:   # let the library report "no injected transitions recorded"
            L.check(L.lib().gmcmc_read_diagnostics(self._h, None, None, None, None))
        la = np.empty((n, self.n_chains), np.float64 if getattr(self, "_discrete", False) else self.dtype)
        acc = np.empty((n, self.n_chains), np.uint8)
        hmc = isinstance(self, HMC)
        pq = np.empty((n, self.n_chains, self.dim), self.dtype) if hmc else None
        pp = np.empty((n, self.n_chains, self.dim), self.dtype) if hmc else None
        L.check(L.lib().gmcmc_read_diagnostics(self._h, L.ptr(la), L.ptr(acc), L.ptr(pq), L.ptr(pp)))
        return {"log_accept": la, "accepted": acc, "prop_q": pq, "prop_p": pp}

    def counters(self):
        c = L.CountersC()
        L.check(L.lib().gmcmc_counters_get(self._h, C.byref(c)))
        return Counters(c.transitions, c.accepts, c.grad_evals, c.divergences, c.step_size, c.kernel_ms, c.launches)


class HMC(_Sampler):
    """≙ hmc::HMC (hmc.rs:75-338) over BatchedGenericHMC (batched_hmc.rs:29-215)."""

    def __init__(self, target, initial_positions, step_size, n_leapfrog, seed=None, ctx=None, chain_offset=0,
                 dtype=None):
        self.ctx = ctx or default_context()
        pos = _as_positions(initial_positions, dtype)
        self.n_chains, self.dim = pos.shape
        self.dtype = self._out_dtype = pos.dtype
        if target.dim is None:
            target.dim = self.dim
        if target.dim != self.dim:
            raise ValueError("target dim %d != position dim %d" % (target.dim, self.dim))
        self.target = target
        self._step_size, self._n_leapfrog = float(step_size), int(n_leapfrog)
        self._th = target._create(self.ctx, self.dtype)
        h = C.c_void_p()
        # the reference seeds from rand::rng() when no seed is given (batched_hmc.rs:80)
        seed = int(np.random.SeedSequence().entropy & 0xFFFFFFFFFFFFFFFF) if seed is None else int(seed)
        L.check(L.lib().gmcmc_hmc_create(self.ctx._h, self._th, C.c_size_t(self.n_chains), C.c_uint64(chain_offset),
                                         L.ptr(pos), C.c_double(step_size), C.c_uint32(n_leapfrog),
                                         C.c_uint64(seed), C.byref(h)))
        self._h = h

    def step_size(self):
        return self.counters().step_size

    def n_leapfrog(self):
        return self._n_leapfrog

    def set_step_size(self, eps):
        L.check(L.lib().gmcmc_set_step_size(self._h, C.c_double(eps)))
        self._step_size = float(eps)

    def set_adaptation(self, mode, target_accept=0.8):
        """mode: 'none' | 'per_chain' | 'pooled' — dual averaging during the discard phase."""
        code = {"none": L.ADAPT_NONE, "per_chain": L.ADAPT_PER_CHAIN, "pooled": L.ADAPT_POOLED}[mode]
        L.check(L.lib().gmcmc_set_adaptation(self._h, code, C.c_double(target_accept)))
        return self


class MetropolisHastings(_Sampler):
    """≙ metropolis_hastings::MetropolisHastings + ChainRunner (metropolis_hastings.rs:90-218,
    core.rs:204-406) with an IsotropicGaussian proposal."""

    def __init__(self, target, proposal, initial_states, ctx=None, chain_offset=0, dtype=None):
        self.ctx = ctx or default_context()
        self._out_dtype = np.dtype(np.float64)   # Trace -> f64, core.rs:34-51
        self.target, self.proposal = target, proposal
        seed = int(np.random.SeedSequence().entropy & 0xFFFFFFFFFFFFFFFF)
        h = C.c_void_p()
        if hasattr(target, "int_kind"):
            # integer-state chains (S = i32): Poisson / Binomial target with the +-1 random-walk proposal
            if not isinstance(proposal, RandomWalkProposal):
                raise TypeError("discrete targets run with RandomWalkProposal")
            pos = np.ascontiguousarray(np.asarray(initial_states), dtype=np.int32)
            if pos.ndim != 2:
                raise ValueError("initial states must be [n_chains, dim]")
            self.n_chains, self.dim = pos.shape
            self.dtype = np.dtype(np.int32)
            self._discrete = True
            p = np.ascontiguousarray(target.params(), np.float64)
            L.check(L.lib().gmcmc_mh_int_create(self.ctx._h, int(target.int_kind), L.ptr(p), C.c_size_t(p.size),
                                                C.c_size_t(self.n_chains), C.c_int(self.dim), C.c_uint64(chain_offset),
                                                L.ptr(pos), C.c_uint64(seed), C.byref(h)))
            self._h = h
            return
        if not isinstance(proposal, IsotropicGaussian):
            raise TypeError("the device path implements the IsotropicGaussian proposal")
        pos = _as_positions(initial_states, dtype)
        self.n_chains, self.dim = pos.shape
        self.dtype = pos.dtype
        if target.dim is None:
            target.dim = self.dim
        self._th = target._create(self.ctx, self.dtype)
        L.check(L.lib().gmcmc_mh_create(self.ctx._h, self._th, C.c_double(proposal.std), C.c_size_t(self.n_chains),
                                        C.c_uint64(chain_offset), L.ptr(pos), C.c_uint64(seed), C.byref(h)))
        self._h = h

    def inject_int(self, steps, ln_u):
        """Test hook for integer-state chains: directions (+1 / -1) [n, chains, dim] and ln u [n, chains]."""
        steps = np.ascontiguousarray(steps, np.int8)
        ln_u = np.ascontiguousarray(ln_u, np.float64)
        n = ln_u.shape[0]
        assert steps.shape == (n, self.n_chains, self.dim) and ln_u.shape == (n, self.n_chains)
        L.check(L.lib().gmcmc_mh_int_inject(self._h, L.ptr(steps), L.ptr(ln_u), C.c_size_t(n)))
        self._n_inj = n

    def seed(self, seed):  # metropolis_hastings.rs:189-197
        return self.set_seed(seed)

    def record(self, n_steps):
        """Test hook (gmcmc_mh_record): the next n_steps transitions of the production 2-D fast kernel record their
        log ratio / decision (diagnostics()) and the draws they used (draws())."""
        L.check(L.lib().gmcmc_mh_record(self._h, C.c_size_t(n_steps)))
        self._n_inj = int(n_steps)

    def draws(self):
        """float32 [n_steps, n_chains, 3]: proposal noise z0, z1 and the accept uniform of the recorded transitions."""
        out = np.empty((self._n_inj, self.n_chains, 3), np.float32)
        L.check(L.lib().gmcmc_mh_read_draws(self._h, L.ptr(out)))
        return out


class ConstantConditional:
    """≙ the ConstantConditional of the reference's Gibbs tests (gibbs.rs:177-186): every coordinate becomes c."""
    cond_kind = 0

    def __init__(self, c):
        self.c = float(c)

    def params(self):
        return [self.c]


class MixtureConditional:
    """≙ MixtureConditional (gibbs.rs:188-245) on the state [x, z]: x | z ~ N(mu_z, sigma_z^2), z | x Bernoulli with
    the posterior weight of mode 1."""
    cond_kind = 1

    def __init__(self, mu0, sigma0, mu1, sigma1, pi0):
        self.mu0, self.sigma0, self.mu1, self.sigma1, self.pi0 = (float(v) for v in (mu0, sigma0, mu1, sigma1, pi0))

    def params(self):
        return [self.mu0, self.sigma0, self.mu1, self.sigma1, self.pi0]


class CustomConditional:
    """A conditional compiled ahead of time into a plugin (csrc/gmcmc_custom_conditional.cuh); the plugin fixes dim."""

    def __init__(self, plugin_path, params=()):
        self.plugin_path = str(plugin_path)
        self._params = [float(v) for v in params]

    def params(self):
        return self._params


class GibbsSampler(_Sampler):
    """≙ gibbs::GibbsSampler (gibbs.rs:107-162) + ChainRunner (core.rs:204-406): `run(n_collect, n_discard)` returns f64
    [n_chains, n_collect, dim]; one transition is one full sweep over the coordinates (gibbs.rs:89-105)."""

    def __init__(self, target, initial_states, ctx=None, chain_offset=0):
        self.ctx = ctx or default_context()
        self._out_dtype = np.dtype(np.float64)
        self.dtype = np.dtype(np.float64)
        self.target = target
        pos = _as_positions(initial_states, np.float64)
        self.n_chains, self.dim = pos.shape
        seed = int(np.random.SeedSequence().entropy & 0xFFFFFFFFFFFFFFFF)
        p = np.ascontiguousarray(target.params(), np.float64)
        h = C.c_void_p()
        if isinstance(target, CustomConditional):
            L.check(L.lib().gmcmc_gibbs_create_custom(self.ctx._h, target.plugin_path.encode(), L.ptr(p) if p.size else None,
                                                      C.c_size_t(p.size), C.c_size_t(self.n_chains), C.c_uint64(chain_offset),
                                                      L.ptr(pos), C.c_uint64(seed), C.byref(h)))
        else:
            L.check(L.lib().gmcmc_gibbs_create(self.ctx._h, int(target.cond_kind), L.ptr(p), C.c_size_t(p.size),
                                               C.c_size_t(self.n_chains), C.c_int(self.dim), C.c_uint64(chain_offset),
                                               L.ptr(pos), C.c_uint64(seed), C.byref(h)))
        self._h = h

    def inject(self, normals, uniforms):
        """Test hook: the first normal / uniform of every (sweep, coordinate), f64 [n, chains, dim] each."""
        normals = np.ascontiguousarray(normals, np.float64)
        uniforms = np.ascontiguousarray(uniforms, np.float64)
        n = normals.shape[0]
        assert normals.shape == (n, self.n_chains, self.dim) and uniforms.shape == normals.shape
        L.check(L.lib().gmcmc_gibbs_inject(self._h, L.ptr(normals), L.ptr(uniforms), C.c_size_t(n)))


class NUTSMassMatrixConfig:
    """≙ NUTSMassMatrixConfig (generic_nuts.rs:40-78): warm-up mass-matrix adaptation of NUTS.  `adaptation` is
    "none", "diagonal" (the reference default) or "dense" (per-chain dense covariance up to dense_max_dim dimensions)."""
    _KINDS = {"none": 0, "diagonal": 1, "dense": 2}

    def __init__(self, adaptation="diagonal", start_buffer=75, end_buffer=50, initial_window=25, regularize=0.05,
                 jitter=1e-6, dense_max_dim=75):
        if adaptation not in self._KINDS:
            raise ValueError("adaptation must be one of %s" % sorted(self._KINDS))
        self.adaptation, self.start_buffer, self.end_buffer = adaptation, int(start_buffer), int(end_buffer)
        self.initial_window, self.regularize, self.jitter = int(initial_window), float(regularize), float(jitter)
        self.dense_max_dim = int(dense_max_dim)

    @classmethod
    def disabled(cls):
        return cls("none", 0, 0, 0, 0.0, 0.0, 0)


class NUTS(_Sampler):
    """≙ nuts::NUTS (nuts.rs:156-304) over GenericNUTS (generic_nuts.rs:370-557).  `mass_matrix` ≙
    GenericNUTS::new_with_mass_matrix (generic_nuts.rs:379-398); None = identity mass (GenericNUTS::new)."""

    def __init__(self, target, initial_positions, target_accept_p, seed=None, ctx=None, chain_offset=0, dtype=None,
                 max_depth=0, init_step_size=-1.0, mass_matrix=None):
        self.ctx = ctx or default_context()
        pos = _as_positions(initial_positions, dtype)
        self.n_chains, self.dim = pos.shape
        self.dtype = self._out_dtype = pos.dtype
        if target.dim is None:
            target.dim = self.dim
        self.target = target
        self._th = target._create(self.ctx, self.dtype)
        h = C.c_void_p()
        seed = int(np.random.SeedSequence().entropy & 0xFFFFFFFFFFFFFFFF) if seed is None else int(seed)
        L.check(L.lib().gmcmc_nuts_create(self.ctx._h, self._th, C.c_size_t(self.n_chains), C.c_uint64(chain_offset),
                                          L.ptr(pos), C.c_double(target_accept_p), C.c_uint32(max_depth),
                                          C.c_double(init_step_size), C.c_uint64(seed), C.byref(h)))
        self._h = h
        if mass_matrix is not None and mass_matrix.adaptation != "none":
            m = mass_matrix
            self._mass_kind = m.adaptation if not (m.adaptation == "dense" and self.dim > m.dense_max_dim) else "none"
            L.check(L.lib().gmcmc_nuts_set_dense_max_dim(self._h, C.c_size_t(m.dense_max_dim)))
            L.check(L.lib().gmcmc_nuts_set_mass_adaptation(self._h, C.c_int(m._KINDS[m.adaptation]),
                                                           C.c_size_t(m.start_buffer), C.c_size_t(m.end_buffer),
                                                           C.c_size_t(m.initial_window), C.c_double(m.regularize),
                                                           C.c_double(m.jitter)))

    def mass_matrix(self):
        """Per-chain inverse mass — [chains, dim] (diagonal) or [chains, dim, dim] (dense) — and the number of warm-up
        updates applied so far."""
        shape = (self.n_chains, self.dim, self.dim) if getattr(self, "_mass_kind", "") == "dense" else (self.n_chains, self.dim)
        inv = np.empty(shape, self.dtype)
        n = C.c_uint64(0)
        L.check(L.lib().gmcmc_nuts_mass_matrix(self._h, L.ptr(inv), C.byref(n)))
        return inv, int(n.value)

    def state(self):
        """Per-chain step sizes, accumulated leapfrogs and (when streams were injected) consumed draws."""
        eps = np.empty(self.n_chains, self.dtype)
        leap = np.empty(self.n_chains, np.int64)
        used = np.zeros((self.n_chains, 3), np.uint64) if getattr(self, "_injected", False) else None
        L.check(L.lib().gmcmc_nuts_state(self._h, L.ptr(eps), L.ptr(leap), L.ptr(used)))
        return {"eps": eps, "leapfrogs": leap, "used": used}

    def inject_streams(self, normals, exp1, unif):
        self._injected = True
        normals = np.ascontiguousarray(normals, np.float64)
        exp1 = np.ascontiguousarray(exp1, np.float64)
        unif = np.ascontiguousarray(unif, np.float64)
        L.check(L.lib().gmcmc_nuts_inject(self._h, L.ptr(normals), C.c_size_t(normals.shape[1]), L.ptr(exp1),
                                          C.c_size_t(exp1.shape[1]), L.ptr(unif), C.c_size_t(unif.shape[1])))


# ------------------------------------------------------------------------------------------------
# initial positions (core.rs:434-475).  The reference draws from Xoshiro256++ + ziggurat (rand 0.9 /
# rand_distr 0.5, not reproducible offline: SURVEY 8c); these use numpy's PCG64 instead.
# ------------------------------------------------------------------------------------------------
def init(n, d, dtype=np.float64):
    return np.random.default_rng().standard_normal((n, d)).astype(dtype)


def init_with_seed(n, d, seed, dtype=np.float64):
    return np.random.default_rng(seed).standard_normal((n, d)).astype(dtype)


def init_det(n, d, dtype=np.float64):
    return init_with_seed(n, d, 42, dtype)
