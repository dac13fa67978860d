"""One 20-transition K1 launch with sample write-out under the bench's timing protocol, piece by piece: which of the L2 flush,
the device-side delay and a cold start costs what (every sample printed, not the best)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C  # noqa: E402
import general_mcmc_b200 as gm  # noqa: E402
from general_mcmc_b200 import _lib as L  # noqa: E402

ctx = gm.default_context()
chains = 65536
q0 = (1.0 + 0.1 * np.random.default_rng(1).standard_normal((chains, 100))).astype(np.float32)
s = gm.HMC(gm.RosenbrockND(100), q0, 0.0119, 32, seed=42, ctx=ctx)
s.reserve(128)
s.run_device(128, 0)
A = torch.randn(4096, 4096, device="cuda")
B = torch.empty_like(A)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
stream = torch.cuda.ExternalStream(ctx.stream())


def sample(mode):
    out = []
    for _ in range(5):
        if "flush" in mode:
            flush.fill_(1)
        torch.cuda.synchronize()
        if "gap" in mode:
            time.sleep(0.05)          # the GPU idles, as it does around the bench's barrier
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            if "sleep" in mode:
                torch.cuda._sleep(400000)
            if "preroll" in mode:     # what bench.py does: read-modify-write passes over the flush buffer
                for _ in range(3):
                    flush.add_(1)
            if "warm" in mode:        # what bench.py does now: all-SM FFMA work from the library itself
                L.check(L.lib().gmcmc_ctx_warm_fp32(ctx._h, C.c_double(4.0)))
            if "busy" in mode:        # ~300 us of all-SM work instead of the one-thread spin
                for _ in range(3):
                    torch.mm(A, A, out=B)
            e0.record(stream)
            s.run_device(20, 0)
            e1.record(stream)
        torch.cuda.synchronize()
        out.append((round(e0.elapsed_time(e1) * 1e3, 1), round(s.counters().kernel_ms * 1e3, 1)))
    print("%-14s (event us, library kernel us): %s" % (mode, out), flush=True)


for mode in ("plain", "gap", "gap+sleep", "gap+flush+sleep", "gap+busy", "gap+flush+busy", "gap+flush+preroll", "gap+flush+warm", "plain"):
    sample(mode)
