"""Multi-GPU acceptance script (run under torchrun with one process per GPU; tests/test_gpu_multi.py launches it):
  1. G-invariance: the chains of a sharded run are bit-identical to the same global chains of a 1-GPU run;
  2. device RunStats over all ranks' chains (NCCL collectives A2 + A3) == RunStats of the gathered samples;
  3. pooled dual averaging (collective A1) adapts every rank to the same step size.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import general_mcmc_b200 as gm  # noqa: E402
from general_mcmc_b200 import dist as gdist  # noqa: E402


def main():
    rank, world, local = gdist.env()
    torch.cuda.set_device(local)
    ctx = gdist.make_context(local)
    C, d, n = 4096, 20, 200
    q0 = (1.0 + 0.05 * np.random.default_rng(0).standard_normal((C, d))).astype(np.float32)
    lo, hi = gm.shard_chains(C, rank, world)

    # 1 + 2: fixed step size
    s = gm.HMC(gm.RosenbrockND(d), q0[lo:hi], 0.01, 8, seed=42, ctx=ctx, chain_offset=lo)
    mine, st = s.run_progress(n, 50)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    full = np.concatenate(gathered)
    ok = True
    if rank == 0:
        solo_ctx = gm.Context(local)
        ref = gm.HMC(gm.RosenbrockND(d), q0, 0.01, 8, seed=42, ctx=solo_ctx)
        ref_samples, ref_st = ref.run_progress(n, 50)
        same = np.array_equal(full, ref_samples)
        print("G-invariance of samples:", same)
        rel = max(abs(st.ess.min / ref_st.ess.min - 1), abs(st.ess.max / ref_st.ess.max - 1),
                  abs(st.rhat.min / ref_st.rhat.min - 1), abs(st.rhat.max / ref_st.rhat.max - 1))
        print("RunStats over ranks vs single GPU: max rel diff %.2e" % rel, st.ess, ref_st.ess)
        ok = ok and same and rel < 1e-4
    # 3: pooled adaptation
    a = gm.HMC(gm.RosenbrockND(d), q0[lo:hi], 0.01, 8, seed=7, ctx=ctx, chain_offset=lo).set_adaptation("pooled", 0.8)
    a.run(10, 150)
    eps = a.counters().step_size
    all_eps = [None] * world
    dist.all_gather_object(all_eps, eps)
    if rank == 0:
        solo = gm.HMC(gm.RosenbrockND(d), q0, 0.01, 8, seed=7, ctx=solo_ctx).set_adaptation("pooled", 0.8)
        solo.run(10, 150)
        e1 = solo.counters().step_size
        print("pooled step size per rank:", all_eps, "single GPU:", e1)
        ok = ok and len(set(all_eps)) == 1 and abs(all_eps[0] / e1 - 1) < 1e-5
        print("MULTIGPU_OK" if ok else "MULTIGPU_FAIL")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
