"""Times K4 (device split R-hat / ESS) alone on a resident [chains, n, p] f32 tensor: tools/stats_bench.py [chains n p]."""
import ctypes as C
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch

import general_mcmc_b200 as gm
from general_mcmc_b200 import _lib as L

chains, n, p = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (65536, 500, 100)
ctx = gm.default_context()
x = torch.randn(chains, n, p, device="cuda", dtype=torch.float32)
torch.cuda.synchronize()
rhat = np.empty(p, np.float32)
ess = np.empty(p, np.float32)
stream = torch.cuda.ExternalStream(ctx.stream_handle()) if hasattr(ctx, "stream_handle") else None
times = []
y = torch.empty(64 << 20, device="cuda")
for it in range(12):
    y.normal_()                       # keeps the clocks up and evicts the L2 between calls
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    L.check(L.lib().gmcmc_split_rhat_ess(ctx._h, C.c_void_p(x.data_ptr()), C.c_size_t(chains), C.c_size_t(n), C.c_size_t(p),
                                         L.F32, 1, L.ptr(rhat), L.ptr(ess)))
    ctx.synchronize()
    times.append((time.perf_counter() - t0) * 1e3)
times = sorted(times[2:])
best = times[0]
print("wall ms: min %.2f median %.2f max %.2f" % (times[0], times[len(times) // 2], times[-1]))
gb = chains * n * p * 4 / 1e9
print("K4 %d x %d x %d: %.2f ms wall (incl. buffer allocation and the host read-back), %.0f GB/s of sample reads; ess mean %.0f (iid: %d), rhat max %.5f"
      % (chains, n, p, best, gb / (best * 1e-3), ess.mean(), chains * n, rhat.max()))

# K6: progress tracker over the same tensor
mx = C.c_float(0)
pa = C.c_float(0)
tt = []
for it in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    L.check(L.lib().gmcmc_tracker_stats(ctx._h, C.c_void_p(x.data_ptr()), C.c_size_t(chains), C.c_size_t(n), C.c_size_t(p),
                                        L.F32, 1, L.ptr(rhat), C.byref(mx), C.byref(pa)))
    ctx.synchronize()
    tt.append((time.perf_counter() - t0) * 1e3)
print("K6 tracker: %.2f ms wall, %.0f GB/s of sample reads; max_rhat %.5f p_accept %.4f" % (min(tt[1:]), gb / (min(tt[1:]) * 1e-3), mx.value, pa.value))
