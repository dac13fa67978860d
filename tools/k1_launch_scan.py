"""Duration of ONE plain K1 launch (fixed step size by value, no adaptation, no sample write-out) versus the number of
transitions in it: separates the per-launch fixed cost from the per-transition cost."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import general_mcmc_b200 as gm  # noqa: E402

ctx = gm.default_context()
WRITE = os.environ.get("K1_SCAN_WRITE", "0") == "1"     # 1: the transitions are collected (sample write-out), as in bench.py
for chains in (32768, 65536):
    q0 = (1.0 + 0.1 * np.random.default_rng(1).standard_normal((chains, 100))).astype(np.float32)
    s = gm.HMC(gm.RosenbrockND(100), q0, 0.015, 32, seed=42, ctx=ctx)
    s.run_device(0, 64)
    if WRITE:
        s.reserve(128)
        s.run_device(128, 0)
    for n in (1, 2, 4, 8, 20, 32, 128):
        best = 1e9
        for _ in range(6):
            if WRITE:
                s.run_device(n, 0)
            else:
                s.run_device(0, n)
            best = min(best, s.counters().kernel_ms)
        print(json.dumps({"write_out": WRITE, "chains": chains, "transitions_per_launch": n, "launch_ms": best, "ms_per_transition": best / n}), flush=True)
    s.close()
