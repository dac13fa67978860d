"""A few single-transition K1 launches (profiling target for the per-launch fixed cost)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import general_mcmc_b200 as gm  # noqa: E402

ctx = gm.default_context()
chains = 65536
q0 = (1.0 + 0.1 * np.random.default_rng(1).standard_normal((chains, 100))).astype(np.float32)
s = gm.HMC(gm.RosenbrockND(100), q0, 0.015, 32, seed=42, ctx=ctx)
for _ in range(6):
    s.run_device(0, 1)
print(s.counters())
