"""general_mcmc_b200 — B200-native (sm_100a) many-chain sampling hot path of general-mcmc.

Host-side mirror of the reference's sampler front-ends (api.py) over the C ABI of include/gmcmc.h
(csrc/ -> libgmcmc.so).  The CUDA library is required: importing the package works without it (so the
build step can run), any call raises if it is missing.
"""
from ._lib import GmcmcError, LIB_PATH, SYMBOLS, lib  # noqa: F401
from .api import (  # noqa: F401
    HMC, NUTS, NUTSMassMatrixConfig, BasicStats, Context, Counters, CustomTarget, DenseGaussian, DiffableGaussian2D, Gaussian2D, GaussianMixture,
    IsotropicGaussian, MetropolisHastings, GibbsSampler, ConstantConditional, MixtureConditional, CustomConditional, PoissonTarget, BinomialTarget, RandomWalkProposal, Rosenbrock2D, RosenbrockND, RunStats, default_context, init, init_det,
    init_with_seed, set_default_context, shard_chains, split_rhat_mean_ess, tracker_stats, build_custom_target, build_custom_conditional)

__version__ = "0.1.0"
