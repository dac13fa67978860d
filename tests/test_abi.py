"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/gmcmc.h declares; compute entry points fail loudly (no CPU fallback) when no GPU is usable."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import general_mcmc_b200 as gm
from general_mcmc_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gmcmc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(gmcmc_[a-z0-9_]+)\s*\(", text))
    return names


def test_header_declares_what_python_binds():
    declared = _declared_symbols()
    assert declared == set(L.SYMBOLS), (declared ^ set(L.SYMBOLS))


def test_library_exports_every_declared_symbol():
    lib = L.lib()
    for name in sorted(_declared_symbols()):
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.gmcmc_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(gm.GmcmcError) as e:
        gm.Context(0)
    assert e.value.status == 2          # GMCMC_ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_null_arguments_are_rejected_not_crashed():
    lib = L.lib()
    assert lib.gmcmc_ctx_create(0, None) == 1             # GMCMC_ERR_INVALID
    assert lib.gmcmc_sampler_destroy(None) == 0
    assert lib.gmcmc_target_destroy(None) == 0
    assert lib.gmcmc_run(None, C.c_size_t(1), C.c_size_t(0), None, 0) == 1
    assert lib.gmcmc_last_error() != b""


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "general_mcmc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inc")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle_lib" not in src and "liboracle" not in src and "gmcmc_oracle" not in src, f


def test_shard_chains_partitions_exactly():
    for n in (1, 7, 65536, 262144, 1000003):
        for world in (1, 2, 3, 8):
            ranges = [gm.shard_chains(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1
