"""Host-side multi-GPU plumbing (one process per GPU, torch.distributed for rendezvous only).

Chains shard contiguously over ranks (api.shard_chains); the data path has no collective.  The tiny
cross-chain reductions (pooled dual averaging, R-hat moments, summed power spectrum) run inside
libgmcmc.so over its own NCCL communicator, whose unique id is exchanged here.
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


def chain_checksums(samples):
    """One 64-bit checksum per chain of a [chains, n, d] sample tensor (sum of the raw bit patterns, wrapping):
    what ranks exchange to verify that a sharded run reproduces the unsharded chains bit for bit."""
    a = np.ascontiguousarray(samples)
    bits = a.view(np.uint32 if a.dtype == np.float32 else np.uint64).reshape(a.shape[0], -1).astype(np.uint64)
    w = (np.arange(bits.shape[1], dtype=np.uint64) * np.uint64(2654435761) + np.uint64(1))
    return (bits * w).sum(axis=1, dtype=np.uint64)


def pin_rank_to_local_cores(local_rank, ranks_on_node):
    """Binds this process to its share of the host cores (contiguous slice of the allowed set) so that the ranks of
    one node do not migrate over each other's cores while they drive PCIe copies; returns the cores kept."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
    except AttributeError:
        return None
    n = max(1, int(ranks_on_node))
    per = max(1, len(allowed) // n)
    mine = allowed[(local_rank % n) * per:(local_rank % n + 1) * per] or allowed
    try:
        os.sched_setaffinity(0, mine)
    except OSError:
        return None
    return mine


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
