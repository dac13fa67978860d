"""K1 scan: steady-state grad-evals/s of the HMC trajectory kernel versus chains per GPU (wave quantisation), and the
cost per transition of the pooled dual-averaging warm-up versus plain sampling.  Run on the GPU box."""
import json
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import general_mcmc_b200 as gm  # noqa: E402

ctx = gm.default_context()
d, L = 100, 32
out = []
for chains in (14208, 16384, 28416, 32768, 65536, 131072, 262144):
    q0 = (1.0 + 0.1 * np.random.default_rng(1).standard_normal((chains, d))).astype(np.float32)
    s = gm.HMC(gm.RosenbrockND(d), q0, 0.015, L, seed=42, ctx=ctx)
    n = 100
    s.reserve(n)
    s.run_device(n, 0)
    best = 1e9
    for _ in range(3):
        s.run_device(n, 0)
        best = min(best, s.counters().kernel_ms)
    rate = chains * n * L / (best * 1e-3)
    # pooled warm-up: 100 transitions, one per launch, DA chain on the side stream
    s.set_adaptation("pooled", 0.8)
    s.run_device(0, 20)
    wbest = 1e9
    for _ in range(3):
        s.run_device(0, 100)
        wbest = min(wbest, s.counters().kernel_ms)
    row = {"chains": chains, "ms_per_transition": best / n, "grad_evals_per_s": rate, "warmup_ms_per_transition": wbest / 100,
           "warmup_cost_ratio": (wbest / 100) / (best / n)}
    print(json.dumps(row), flush=True)
    out.append(row)
    s.close()
