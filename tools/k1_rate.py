"""Steady-state K1 rate at 65,536 chains (d = 100, L = 32, f32) for the library selected by GMCMC_LIB."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import general_mcmc_b200 as gm  # noqa: E402

ctx = gm.default_context()
chains = 65536
q0 = (1.0 + 0.1 * np.random.default_rng(1).standard_normal((chains, 100))).astype(np.float32)
s = gm.HMC(gm.RosenbrockND(100), q0, 0.015, 32, seed=42, ctx=ctx)
s.reserve(200)
s.run_device(200, 0)
best = 1e9
for _ in range(4):
    s.run_device(200, 0)
    best = min(best, s.counters().kernel_ms)
print(os.environ.get("GMCMC_LIB", "default"), "grad-evals/s %.4e  ms/transition %.5f" % (chains * 200 * 32 / (best * 1e-3), best / 200))
