"""Cost per warm-up transition of the pooled dual-averaging path versus the window length (fixed windows):
time = (n / W) * (F + W * v)  ->  per-launch fixed cost F and per-transition cost v with a device-resident step size."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, json, os
import numpy as np
sys.path.insert(0, %r)
import general_mcmc_b200 as gm
ctx = gm.default_context()
chains = int(sys.argv[1])
q0 = (1.0 + 0.1 * np.random.default_rng(1).standard_normal((chains, 100))).astype(np.float32)
s = gm.HMC(gm.RosenbrockND(100), q0, 0.015, 32, seed=42, ctx=ctx)
s.run_device(0, 64)
base = 1e9
for _ in range(3):
    s.run_device(0, 256); base = min(base, s.counters().kernel_ms)
s.set_adaptation("pooled", 0.8)
s.run_device(0, 64)
best = 1e9
for _ in range(3):
    s.run_device(0, 256); best = min(best, s.counters().kernel_ms)
print(json.dumps({"chains": chains, "window": os.environ.get("GMCMC_POOLED_WINDOW"), "plain_ms_per_transition": base / 256,
                  "warmup_ms_per_transition": best / 256, "ratio": best / base, "eps": s.counters().step_size}))
''' % ROOT
for chains in (32768, 65536):
    for w in (1, 2, 4, 8, 16, 64):
        env = dict(os.environ, GMCMC_POOLED_WINDOW=str(w), GMCMC_POOLED_FIXED="1")
        r = subprocess.run([sys.executable, "-c", CODE, str(chains)], env=env, capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr[-500:], flush=True)
