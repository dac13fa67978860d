"""World-size-2 CPU tests (gloo) of the host-side multi-GPU plumbing: the NCCL unique-id exchange, the
chain sharding, the max-over-ranks timing reduction and the cross-rank combination of R-hat moments
(collective A2 of SURVEY 8e) follow the same code the GPU ranks run, with gloo standing in for NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- collective A2 restated on the host (what stats_accumulate / stats_finalize do on the device): test-side
# reference of the moment combination, so the N > 1 logic is checkable on CPU ----
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


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import general_mcmc_b200 as gm
    from general_mcmc_b200 import dist as gdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. unique-id exchange: rank 0's 128 bytes reach every rank
        nid = gdist.broadcast_bytes(bytes(range(128)) if rank == 0 else None, 128, src=0)
        assert nid == bytes(range(128))
        # 2. sharding: contiguous, disjoint, covering
        C = 1001
        lo, hi = gm.shard_chains(C, rank, world)
        sizes = gdist.all_gather_object((lo, hi))
        assert sizes[0][0] == 0 and sizes[-1][1] == C and sizes[0][1] == sizes[1][0]
        # 3. timing: max over ranks
        t = gdist.max_over_ranks(1.0 + rank)
        assert t == float(world)
        # 4. split R-hat from per-rank moment partials == single-process result (collective A2)
        rng = np.random.default_rng(0)
        x = rng.standard_normal((C, 40, 3)).astype(np.float32) + np.arange(3, dtype=np.float32)
        part = rhat_moment_partials(x[lo:hi])
        tot = gdist.all_reduce_sum(part)
        rhat = rhat_from_moments(tot, n=40)
        full = rhat_from_moments(rhat_moment_partials(x), n=40)
        assert np.allclose(rhat, full, rtol=1e-6)
        # 5. G-invariance bookkeeping used by bench.py / tests/multigpu_check.py: per-chain checksums of the shards,
        # gathered in rank order, equal the checksums of the unsharded tensor
        mine = gdist.chain_checksums(x[lo:hi])
        allsums = np.concatenate(gdist.all_gather_object(mine))
        assert np.array_equal(allsums, gdist.chain_checksums(x))
        y = x.copy(); y[7, 3, 1] = np.nextafter(y[7, 3, 1], np.float32(10))
        assert (gdist.chain_checksums(y) != gdist.chain_checksums(x)).sum() == 1
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_host_plumbing_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
