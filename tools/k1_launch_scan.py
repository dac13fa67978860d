"""Duration of ONE plain K1 launch (fixed step size by value, no adaptation, no sample write-out) versus the number of
transitions in it: separates the per-launch fixed cost from the per-transition cost."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import general_mcmc_b200 as gm  # noqa: E402

ctx = gm.default_context()
for chains in (32768, 65536):
    q0 = (1.0 + 0.1 * np.random.default_rng(1).standard_normal((chains, 100))).astype(np.float32)
    s = gm.HMC(gm.RosenbrockND(100), q0, 0.015, 32, seed=42, ctx=ctx)
    s.run_device(0, 64)
    for n in (1, 2, 4, 8, 32, 128):
        best = 1e9
        for _ in range(6):
            s.run_device(0, n)
            best = min(best, s.counters().kernel_ms)
        print(json.dumps({"chains": chains, "transitions_per_launch": n, "launch_ms": best, "ms_per_transition": best / n}), flush=True)
    s.close()
