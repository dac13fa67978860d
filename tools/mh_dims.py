"""K2 generic kernel (d > 2): chain-steps/s and HBM write rate of mh_run_kernel for an isotropic Gaussian target at
d = 2 (the fast 2-D kernel, for comparison), 4, 8, 16, 32:  python tools/mh_dims.py [chains] [steps]"""
import sys

sys.path.insert(0, ".")
import numpy as np
import torch

import general_mcmc_b200 as gm

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
ctx = gm.default_context()
for d in (2, 4, 8, 16, 32):
    n = max(64, min(steps, int(12e9 / (chains * d * 8))))
    x0 = np.random.default_rng(d).standard_normal((chains, d))
    s = gm.MetropolisHastings(gm.IsotropicGaussian(1.0, d), gm.IsotropicGaussian(2.4 / np.sqrt(d), d), x0, ctx=ctx).seed(3)
    s.reserve(n)
    s.run_device(n, 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(3):
        ctx.synchronize()
        torch.cuda.synchronize()
        e0.record()
        s.run_device(n, 0)
        ctx.synchronize()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    rate = chains * n / (best * 1e-3)
    print("d=%2d chains=%d steps=%d: %.2f ms, %.3e chain-steps/s, %.0f GB/s written (%.2f of 6543), accept %.3f"
          % (d, chains, n, best, rate, rate * d * 8 / 1e9, rate * d * 8 / 1e9 / 6543, s.counters().accept_rate))
    s.close()
