"""Host-side multi-GPU plumbing (one process per GPU, torch.distributed for rendezvous only).

Chains shard contiguously over ranks (api.shard_chains); the data path has no collective.  The tiny
cross-chain reductions (pooled dual averaging, R-hat moments, summed power spectrum) run inside
libgmcmc.so over its own NCCL communicator, whose unique id is exchanged here.  The numpy helpers at the
bottom restate the moment combination of collective A2 so the N > 1 logic is testable on CPU (gloo).
"""
import os

import numpy as np


def env():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def broadcast_bytes(payload, n, src=0):
    """Broadcasts `n` bytes from rank `src` over the default process group (any backend)."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(n, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        t = torch.tensor(list(payload), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src)
    return bytes(t.cpu().tolist())


def all_gather_object(obj):
    import torch.distributed as dist
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def max_over_ranks(x):
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_reduce_sum(a):
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.as_tensor(np.ascontiguousarray(a, np.float64), device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def make_context(device=None):
    """Context for this rank: initialises torch.distributed (NCCL) when WORLD_SIZE > 1 and hands the NCCL
    unique id of libgmcmc's own communicator to every rank."""
    from . import api
    rank, world, local = env()
    device = local if device is None else device
    if world == 1:
        return api.Context(device)
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        torch.cuda.set_device(device)
        dist.init_process_group("nccl", device_id=torch.device("cuda", device))
    nid = broadcast_bytes(api.Context.nccl_unique_id() if rank == 0 else None, 128, src=0)
    return api.Context(device, rank, world, nid)


# ---- collective A2 restated on the host (what stats_accumulate / stats_finalize do on the device) ----
def rhat_moment_partials(samples):
    """[sum of split-chain means, sum of their squares, sum of within variances, chains] per parameter,
    stats.rs:419-504 with the chain split (splitcat)."""
    s = np.asarray(samples, np.float64)
    c, n, p = s.shape
    half = n // 2
    halves = np.concatenate([s[:, :half], s[:, n - half:]], axis=0)
    m = halves.mean(axis=1)
    w = ((halves - m[:, None, :]) ** 2).mean(axis=1)
    return np.stack([m.sum(0), (m * m).sum(0), w.sum(0), np.full(p, float(c))])


def rhat_from_moments(tot, n):
    sm, sm2, sw, c = tot
    half = n // 2
    c2 = 2.0 * c
    om = sm / c2
    b = (sm2 - c2 * om * om) * (half / (c2 - 1.0))
    w = sw / c2
    v = (half - 1.0) / half * w + b / half
    return np.sqrt(w / v)          # reference orientation, stats.rs:452-454
