"""ctypes binding of libgmcmc.so (the C ABI in include/gmcmc.h).

There is no CPU fallback: if the shared library is missing or no sm_100 device is usable, the
calls raise.  Nothing in this package imports or executes oracle/.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GMCMC_LIB") or os.path.join(_HERE, "libgmcmc.so")

F32, F64 = 0, 1
MATH_FAST, MATH_EXACT = 0, 1
ADAPT_NONE, ADAPT_PER_CHAIN, ADAPT_POOLED = 0, 1, 2
(TARGET_ISO_GAUSS, TARGET_GAUSS2D, TARGET_DIFF_GAUSS2D, TARGET_DENSE_GAUSS, TARGET_ROSENBROCK2D,
 TARGET_ROSENBROCK_ND, TARGET_GAUSS_MIXTURE) = range(7)

# every entry point include/gmcmc.h declares (checked by tests/test_abi.py against the header)
SYMBOLS = [
    "gmcmc_ctx_create", "gmcmc_nccl_unique_id", "gmcmc_ctx_create_dist", "gmcmc_ctx_destroy",
    "gmcmc_ctx_synchronize", "gmcmc_ctx_stream", "gmcmc_ctx_all_reduce_f64", "gmcmc_host_alloc",
    "gmcmc_host_free", "gmcmc_measure_fp32_peak", "gmcmc_ctx_warm_fp32", "gmcmc_target_create", "gmcmc_target_create_custom", "gmcmc_target_destroy", "gmcmc_target_logp_grad",
    "gmcmc_hmc_create", "gmcmc_mh_create", "gmcmc_mh_int_create", "gmcmc_mh_int_inject", "gmcmc_gibbs_create", "gmcmc_gibbs_create_custom", "gmcmc_gibbs_inject", "gmcmc_nuts_create", "gmcmc_sampler_destroy", "gmcmc_set_seed",
    "gmcmc_set_math_mode", "gmcmc_set_adaptation", "gmcmc_set_step_size", "gmcmc_inject", "gmcmc_mh_record", "gmcmc_mh_read_draws",
    "gmcmc_nuts_inject", "gmcmc_nuts_state", "gmcmc_nuts_set_mass_adaptation", "gmcmc_nuts_set_dense_max_dim", "gmcmc_nuts_mass_matrix", "gmcmc_read_diagnostics", "gmcmc_step", "gmcmc_run", "gmcmc_run_device", "gmcmc_reserve_samples",
    "gmcmc_run_stats", "gmcmc_positions", "gmcmc_set_positions", "gmcmc_counters_get",
    "gmcmc_sampler_info", "gmcmc_split_rhat_ess", "gmcmc_run_stats_from", "gmcmc_tracker_stats", "gmcmc_export_columns", "gmcmc_philox_blocks",
    "gmcmc_last_error", "gmcmc_version",
]


class BasicStatsC(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("min", "median", "max", "mean", "std")]


class RunStatsC(C.Structure):
    _fields_ = [("ess", BasicStatsC), ("rhat", BasicStatsC), ("rhat_std", BasicStatsC)]


class CountersC(C.Structure):
    _fields_ = [("transitions", C.c_uint64), ("accepts", C.c_uint64), ("grad_evals", C.c_uint64),
                ("divergences", C.c_uint64), ("step_size", C.c_double), ("kernel_ms", C.c_double),
                ("launches", C.c_uint64)]


class GmcmcError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("gmcmc status %d: %s" % (status, message))
        self.status = status


_lib = None


def lib():
    """Loads libgmcmc.so once.  Raises if it has not been built (python -m general_mcmc_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: build it with `python general_mcmc_b200/build.py` "
                "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.gmcmc_last_error.restype = C.c_char_p
        L.gmcmc_version.restype = C.c_char_p
        for name in SYMBOLS:
            fn = getattr(L, name)
            if name not in ("gmcmc_last_error", "gmcmc_version"):
                fn.restype = C.c_int
        _lib = L
    return _lib


def check(status):
    if status != 0:
        raise GmcmcError(status, lib().gmcmc_last_error().decode("utf-8", "replace"))


def np_dtype(code):
    return np.float32 if code == F32 else np.float64


def dtype_code(dt):
    dt = np.dtype(dt)
    if dt == np.float32:
        return F32
    if dt == np.float64:
        return F64
    raise TypeError("only float32 and float64 are supported, got %s" % dt)


def ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None
