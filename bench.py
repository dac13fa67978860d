#!/usr/bin/env python
"""bench.py — benchmark of the many-chain sampling hot path (BASELINE.json metric: leapfrog grad-evals/sec and
min-ESS/sec at 65,536+ chains on 1/2/4/8 B200 next to the host CPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Default run = the headline line plus every other BASELINE config nested in the same JSON line:

  headline        config 4, per-GPU shard (weak scaling): batched HMC on the 100-D Rosenbrock target, 65,536 chains per
                  GPU, L = 32, f32, fixed step size after a pooled dual-averaging warm-up.  A "step" is one HMC transition
                  of every chain (L gradient evaluations per chain) including the [chains, samples, dim] write-out;
                  `value` = chains x steps x L / device time, inputs resident in HBM.  `e2e` is the same metric through
                  the host-buffer C-ABI call gmcmc_run (H2D of the positions + D2H of the samples inside the timed
                  region); `e2e_stats_only` (gmcmc_run_stats without samples ≙ run_progress) and `e2e_device`
                  (gmcmc_run_device ≙ run_positions) are the reference's other two contracts.
  workloads       mh_gauss2d (config 2), hmc_dense (config 3), nuts_mixture (config 5): value / roofline / e2e /
                  cpu_baseline each, short fixed-size legs.
  cfg4_strong     config 4 as BASELINE.json states it: 262,144 chains SHARDED over the N ranks (strong scaling), 200
                  pooled dual-averaging warm-up transitions (NCCL all-reduce of the acceptance statistic every
                  transition) + 1,000 draws inside the timed region, device RunStats (ESS / split R-hat over all ranks'
                  chains, NCCL A2 + A3) after it.
  g_invariant     a small sharded run reproduces the unsharded chains bit for bit (per-chain checksums).

`--workload NAME` prints that workload alone as the line (hmc_rosenbrock | mh_gauss2d | hmc_dense | nuts_mixture).
One process per GPU; under torchrun the ranks shard the chains; timing = max over ranks of the CUDA-event time between
two barriers.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHAINS_PER_GPU = 65536
STRONG_CHAINS = 262144
DIM = 100
N_LEAPFROG = 32
STEP_SIZE = 0.01
TRANSITIONS_PER_LAUNCH = 100
E2E_TRANSITIONS = 16
E2E_VARIANT_TRANSITIONS = 128   # draws per call of the stats-only / device-resident end-to-end contracts
# Algorithmic FP32 work per gradient evaluation per chain: 6 FMAs per coordinate (t_i, 400 t_i - 2, x_i(.) + 2,
# - 200 t_{i-1} + (.), p += eps g, q += eps p) = 12 flop x d.  SURVEY 8(d)'s 21 d counted mul and add separately and a
# log-density per leapfrog; the log density is only needed at the two trajectory ends (kernel K1).
FLOP_PER_GRAD_EVAL = 12 * DIM
BYTES_PER_STEP_PER_CHAIN = DIM * 4   # sample write-out, f32


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        d["source"] = "measured (MEASURED_PEAKS.json)"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            time.sleep(0.15)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def window(self, t0, t1):
        rows = [l for (t, l) in self.lines if t0 - 0.05 <= t <= t1 + 0.05] or [l for (_, l) in self.lines[-3:]]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in rows:
            f = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except Exception:
                continue
            for k, name in enumerate(names):
                if len(f) > 4 + k and f[4 + k].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if not self.proc:
            return
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def init_positions(seed, n_chains, dim, dtype=np.float32):
    rng = np.random.default_rng(1234 + seed)
    return (1.0 + 0.1 * rng.standard_normal((n_chains, dim))).astype(dtype)


def dense_target(gm, dd):
    rng = np.random.default_rng(0)
    qmat, _ = np.linalg.qr(rng.standard_normal((dd, dd)))
    lam = np.logspace(-1, 1, dd)
    prec = (qmat / lam) @ qmat.T          # P = Q diag(1/lambda) Q^T  
    return gm.DenseGaussian(np.zeros(dd), precision=prec)


def mixture_params(d=DIM, K=4):
    mu = np.stack([(k - 1.5) * (2.0 / np.sqrt(d)) * np.ones(d) for k in range(K)])
    return np.full(K, 1.0 / K), mu


# ------------------------------------------------------------------------------------------------
# CPU legs (oracle "port": the C++ restatement of the reference algorithm, all host threads)
# ------------------------------------------------------------------------------------------------
def host_threads():
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1 to its workers: ignore that)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    oracle_lib.build()
    return oracle_lib


def cpu_hmc_rate(target_seconds=12.0, chains=None):
    O = _oracle()
    threads = host_threads()
    chains = chains or max(threads * 64, 2048)
    q0 = init_positions(0, chains, DIM)
    secs, _, _ = O.hmc_bench(O.ROSENBROCK_ND, [], q0, STEP_SIZE, N_LEAPFROG, 1, seed=1, threads=threads)
    rate1 = chains * N_LEAPFROG / max(secs, 1e-9)
    n_steps = int(max(1, min(200000, target_seconds * rate1 / (chains * N_LEAPFROG))))
    secs, _, _ = O.hmc_bench(O.ROSENBROCK_ND, [], q0, STEP_SIZE, N_LEAPFROG, n_steps, seed=2, threads=threads)
    rate = chains * n_steps * N_LEAPFROG / secs
    return rate, threads, "%d chains x %d transitions x L=%d, d=%d, f32 (%.1f s)" % (chains, n_steps, N_LEAPFROG, DIM, secs)


def cpu_mh_rate(target_seconds=12.0):
    O = _oracle()
    threads = host_threads()
    chains = max(threads * 256, 8192)
    x0 = np.random.default_rng(0).standard_normal((chains, 2))
    params = [0.0, 0.0, 1.0, 0.0, 0.0, 1.0]
    secs, _, _ = O.mh_bench(O.GAUSS2D, params, x0, 1.0, 20, seed=1, threads=threads, keep_samples=True)
    rate1 = chains * 20 / max(secs, 1e-9)
    n_steps = int(max(1, min(1000, target_seconds * rate1 / chains)))
    secs, _, _ = O.mh_bench(O.GAUSS2D, params, x0, 1.0, n_steps, seed=2, threads=threads, keep_samples=True)
    return chains * n_steps / secs, threads, "%d chains x %d steps, f64, samples kept (%.1f s)" % (chains, n_steps, secs)


def cpu_dense_rate(target_seconds=10.0, dd=1000):
    import general_mcmc_b200 as gm
    O = _oracle()
    threads = host_threads()
    tgt = dense_target(gm, dd)
    chains = max(threads * 2, 16)
    q0 = np.random.default_rng(200).standard_normal((chains, dd)).astype(np.float32)
    params = tgt.params()
    secs, _, _ = O.hmc_bench(O.DENSE_GAUSS, params, q0, 0.05, N_LEAPFROG, 1, seed=1, threads=threads)
    rate1 = chains * N_LEAPFROG / max(secs, 1e-9)
    n_steps = int(max(1, min(50, target_seconds * rate1 / (chains * N_LEAPFROG))))
    secs, _, _ = O.hmc_bench(O.DENSE_GAUSS, params, q0, 0.05, N_LEAPFROG, n_steps, seed=2, threads=threads)
    return (chains * n_steps * N_LEAPFROG / secs, threads,
            "%d chains x %d transitions x L=%d, d=%d dense Gaussian, f32 (%.1f s)" % (chains, n_steps, N_LEAPFROG, dd, secs))


def cpu_nuts_rate(target_seconds=10.0):
    """The recursive NUTS restatement (generic_nuts.rs:755-925) on the config-5 mixture: leapfrogs actually taken / s."""
    O = _oracle()
    threads = host_threads()
    w, mu = mixture_params()
    params = np.concatenate([[4, 1.0], w, mu.ravel()])
    chains = max(threads * 32, 256)
    rng = np.random.default_rng(7)
    q0 = rng.standard_normal((chains, DIM)).astype(np.float32)

    def run(n_collect, n_discard):
        n = n_collect + n_discard
        # uniforms: one per doubling + one per merged subtree; 96 per transition covers depth-6 trees on average (an
        # exhausted stream keeps returning 0.75: harmless for a timing run)
        streams = (rng.standard_normal((chains, DIM * (n + 2))), rng.exponential(size=(chains, n + 2)),
                   rng.random((chains, 96 * (n + 1))))
        O.set_threads(threads)
        t0 = time.perf_counter()
        r = O.nuts_run(O.GAUSS_MIXTURE, params, q0, 0.8, 10, -1.0, n_collect, n_discard, *streams, fast=True)
        return time.perf_counter() - t0, float(r["leapfrogs"].sum())
    secs, leaps = run(2, 6)
    n = int(max(8, min(400, target_seconds / max(secs, 1e-6) * 8)))
    secs, leaps = run(n // 2, n - n // 2)
    return leaps / secs, threads, "%d chains x %d transitions (half warm-up), d=%d mixture, depth <= 10, f32 (%.1f s)" % (
        chains, n - 1, DIM, secs)


def run_reference(args, rank, world):
    if rank != 0:
        return
    wl = args.workload or "hmc_rosenbrock"
    budget = min(60.0, max(5.0, 0.02 * args.steps))
    if wl == "mh_gauss2d":
        rate, threads, sample = cpu_mh_rate(30.0)
        metric, unit, dtype = "mh_chain_steps_per_sec", "chain-steps/s", "f64"
        cfg = {"workload": WORKLOAD_NAMES["mh_gauss2d"] % (args.chains or 1048576)}
    elif wl == "hmc_dense":
        rate, threads, sample = cpu_dense_rate(budget)
        metric, unit, dtype = "leapfrog_grad_evals_per_sec", "grad-evals/s", "f32"
        cfg = {"workload": WORKLOAD_NAMES["hmc_dense"] % (args.dim or 1000, args.chains or 65536, N_LEAPFROG)}
    elif wl == "nuts_mixture":
        rate, threads, sample = cpu_nuts_rate(budget)
        metric, unit, dtype = "leapfrog_grad_evals_per_sec", "grad-evals/s", "f32"
        cfg = {"workload": WORKLOAD_NAMES["nuts_mixture"] % (DIM, args.chains or 65536)}
    else:
        rate, threads, sample = cpu_hmc_rate(budget)
        metric, unit, dtype = "leapfrog_grad_evals_per_sec", "grad-evals/s", "f32"
        cfg = {"workload": WORKLOAD_NAMES["hmc_rosenbrock"] % (DIM, args.chains or CHAINS_PER_GPU, N_LEAPFROG)}
    line = {"impl": "reference", "metric": metric, "value": rate, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": rate, "unit": unit, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the Rust reference cannot be built here (no cargo); this is the C++ restatement of its algorithm "
                    "(oracle/, -O3, OpenMP over chains like rayon core.rs:221-225) on all host threads"}
    print(json.dumps(line), flush=True)


WORKLOAD_NAMES = {
    "hmc_rosenbrock": "cfg4 shard: batched HMC, RosenbrockND d=%d, %d chains/GPU, L=%d, f32, pooled dual-averaging warm-up then fixed step",
    "mh_gauss2d": "cfg2: batched MH, Gaussian2D target, IsotropicGaussian proposal, %d chains/GPU, f64 state and output",
    "hmc_dense": "cfg3: batched HMC, dense-covariance Gaussian d=%d, %d chains/GPU, L=%d, eps=0.05, f32, tcgen05 gradient GEMM",
    "nuts_mixture": "cfg5: NUTS, 4-component isotropic Gaussian mixture d=%d, %d chains/GPU, max depth 10, target accept 0.8, f32",
}

# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of each workload's dominant kernel, from the `ncu --set full`
# captures summarised under profiles/ (same launch shapes as the bench legs).  Constants from a capture, not measured in
# the run that prints them: `traffic_source` says which capture.
DEFAULT_CHAINS = {"hmc_rosenbrock": 65536, "mh_gauss2d": 1048576, "hmc_dense": 65536, "nuts_mixture": 65536}
NCU_TRAFFIC = {
    "hmc_rosenbrock": {"bytes": 27.2e6 + 2.592e9, "launch": "65,536 chains x 100 transitions", "source": "profiles/r1_hmc_run_kernel_full.txt"},
    "mh_gauss2d": {"bytes": 17.1e6 + 16.735e9, "launch": "1,048,576 chains x 1000 steps", "source": "profiles/r1_mh_run2_kernel_full.txt"},
    "hmc_dense": {"bytes": 0.666e9 + 0.479e9, "launch": "one persistent-schedule gradient GEMM of 65,536 chains, d = 1000", "source": "profiles/r1_dense_gemm_kick_full.txt"},
    "nuts_mixture": {"bytes": 47.5e6 + 5.380e9, "launch": "65,536 chains x 200 transitions", "source": "profiles/r1_nuts_run_kernel_full.txt"},
}
_TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_traffic.json")
if os.path.exists(_TRAFFIC_FILE):
    try:
        with open(_TRAFFIC_FILE) as _f:
            NCU_TRAFFIC.update(json.load(_f))
    except Exception:
        pass


def traffic_fields(wl, chains, default_shape=True):
    t = NCU_TRAFFIC[wl]
    ok = default_shape and chains == DEFAULT_CHAINS[wl]
    return {"traffic": t["bytes"] if ok else None,
            "traffic_source": "ncu capture %s (%s): a constant from that capture, not measured in this run" % (t["source"], t["launch"])}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class Env:
    def __init__(self, rank, world, local):
        import torch
        import torch.distributed as dist
        import general_mcmc_b200 as gm
        from general_mcmc_b200 import _lib as L
        from general_mcmc_b200 import dist as gdist
        self.torch, self.dist, self.gm, self.L, self.gdist = torch, dist, gm, L, gdist
        self.rank, self.world, self.local = rank, world, local
        if world > 1:
            self.cores = gdist.pin_rank_to_local_cores(local, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
        else:
            self.cores = None
        torch.cuda.set_device(local)
        self.ctx = gdist.make_context(local)   # world > 1: torch.distributed (NCCL) rendezvous + libgmcmc's own communicator
        self.lib = L.lib()
        self.pk = peaks()
        self.stream = torch.cuda.ExternalStream(self.ctx.stream())
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
        self.clocks = ClockSampler(local)
        if rank == 0:
            self.clocks.start()
        self._fp32_peak = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()
        self.ctx.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        return float(self.ctx.all_reduce([x])[0])

    def fp32_peak(self):
        if self._fp32_peak is None:
            v = C.c_double(0)
            self.L.check(self.lib.gmcmc_measure_fp32_peak(self.ctx._h, C.byref(v)))
            self._fp32_peak = v.value
        return self._fp32_peak

    def timed(self, fn, marks=0):
        """L2 flush, barrier, CUDA events on the library's stream around fn(), barrier; max over ranks (ms).  With
        marks > 0, fn(mark) may call mark() to drop intermediate events; returns (total, [segment ms...])."""
        torch = self.torch
        self.flush.fill_(1)
        self.barrier()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(2 + marks)]
        used = [0]

        def mark():
            used[0] += 1
            evs[used[0]].record(self.stream)
        t0 = time.time()
        with torch.cuda.stream(self.stream):
            # device-side pre-roll ahead of the first event, on the library's stream: ~4 ms of all-SM FP32 FMA work
            # (gmcmc_ctx_warm_fp32, no synchronisation).  (1) The host enqueues the event and the first launch while the
            # GPU is still working, so the region starts with the kernel already queued.  (2) The barrier above leaves the
            # GPU idle, and a B200 runs the first millisecond after an idle gap 3 % slow: a 20-transition K1 launch takes
            # 1,087 us after a gap (or after a one-thread spin, or after memory-only work), 1,052 us after all-SM compute
            # (tools/k1_cold_launch.py, profiles/r2_k1_cold_launch.txt) — the state W warm-up steps are meant to establish
            # and the barrier undoes.  The L2 flush above stands: the pre-roll touches no memory.
            self.L.check(self.lib.gmcmc_ctx_warm_fp32(self.ctx._h, C.c_double(4.0)))
            evs[0].record(self.stream)
            if marks:
                fn(mark)
            else:
                fn()
            evs[-1].record(self.stream)
        self.barrier()
        t1 = time.time()
        total = self.max_over_ranks(evs[0].elapsed_time(evs[-1]))
        self.last_window = (t0, t1)
        if not marks:
            return total
        pts = [evs[0]] + evs[1:1 + used[0]] + [evs[-1]]
        return total, [self.max_over_ranks(pts[i].elapsed_time(pts[i + 1])) for i in range(len(pts) - 1)]

    def clocks_now(self):
        return self.clocks.window(*self.last_window) if self.rank == 0 else None

    def pinned(self, shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        self.L.check(self.lib.gmcmc_host_alloc(C.c_size_t(n), C.byref(p)))
        buf = (C.c_char * n).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape), p

    def free_pinned(self, p):
        self.lib.gmcmc_host_free(p)


def e2e_leg(env, s, chains, dim, out_dtype, e2e_T, calls, units_per_call=None, counter_units=False, variants=False):
    """The metric end to end through the host-buffer C ABI: gmcmc_set_positions (H2D from pinned memory) + gmcmc_run
    (samples D2H into pinned memory) per call, wall clock, max over ranks."""
    L, lib, ctx = env.L, env.lib, env.ctx
    host_out, host_ptr = env.pinned((chains, e2e_T, dim), out_dtype)
    init_host, init_ptr = env.pinned((chains, dim), s.dtype)
    init_host[...] = s.positions()

    def measure(call):
        for _ in range(2):
            L.check(lib.gmcmc_set_positions(s._h, L.ptr(init_host)))
            call()
        env.barrier()
        c0 = s.counters()
        t0 = time.perf_counter()
        for _ in range(calls):
            L.check(lib.gmcmc_set_positions(s._h, L.ptr(init_host)))
            call()
        ctx.synchronize()
        secs = env.max_over_ranks(time.perf_counter() - t0)
        if counter_units:
            units = env.sum_over_ranks(s.counters().grad_evals - c0.grad_evals)
        else:
            units = units_per_call * calls * env.world
        return units / secs, secs

    def measure_units(call, n_calls, units_call):
        nonlocal calls, units_per_call
        keep = (calls, units_per_call)
        calls, units_per_call = n_calls, units_call
        try:
            return measure(call)
        finally:
            calls, units_per_call = keep

    value, secs = measure(lambda: s.run(e2e_T, 0, out=host_out))
    h2d = init_host.nbytes / e2e_T
    d2h = host_out.nbytes / e2e_T
    out = {"value": value, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "calls": calls,
           "transitions_per_call": e2e_T, "host_memory": "pinned", "pcie_gbs": (h2d + d2h) * e2e_T * calls / secs / 1e9,
           "note": "gmcmc_set_positions + gmcmc_run into host memory: the [chains, samples, dim] tensor crosses PCIe every "
                   "call, which bounds this figure"}
    extra = {}
    if variants:
        # the two contracts that do not move the sample tensor are measured at a call length a progress display or a
        # run_positions user would use (E2E_VARIANT_TRANSITIONS draws per call; RunStats of 16 draws is not a use case)
        vT = E2E_VARIANT_TRANSITIONS
        v_units = units_per_call / e2e_T * vT
        s.reserve(vT)
        st = L.RunStatsC()
        v2, _ = measure_units(lambda: L.check(lib.gmcmc_run_stats(s._h, C.c_size_t(vT), C.c_size_t(0), None, L.dtype_code(out_dtype),
                                                                   C.byref(st))), 3, v_units)
        extra["e2e_stats_only"] = {"value": v2, "h2d_bytes_per_step": init_host.nbytes / vT, "d2h_bytes_per_step": 60.0 / vT,
                                   "transitions_per_call": vT, "calls": 3,
                                   "api": "gmcmc_set_positions + gmcmc_run_stats(out = NULL) ≙ run_progress: RunStats only "
                                          "(device ESS / split R-hat over all ranks' chains) crosses PCIe"}
        pos = np.empty((chains, dim), s.dtype)

        def dev_call():
            s.run_device(vT, 0)
            L.check(lib.gmcmc_positions(s._h, L.ptr(pos)))
        v3, _ = measure_units(dev_call, 3, v_units)
        extra["e2e_device"] = {"value": v3, "h2d_bytes_per_step": init_host.nbytes / vT, "d2h_bytes_per_step": pos.nbytes / vT,
                               "transitions_per_call": vT, "calls": 3,
                               "api": "gmcmc_set_positions + gmcmc_run_device + gmcmc_positions ≙ run_positions "
                                      "(batched_hmc.rs:115-123): samples stay on the GPU, final positions come back"}
    env.free_pinned(host_ptr)
    env.free_pinned(init_ptr)
    return out, extra


def launches_for(k, per_launch):
    full, rem = divmod(k, per_launch)
    return [per_launch] * full + ([rem] if rem else [])


def leg_hmc_rosenbrock(env, steps, warmup, chains, want_ess=True, want_cpu=True, variants=True):
    gm, L, lib, ctx = env.gm, env.L, env.lib, env.ctx
    rank, world = env.rank, env.world
    per_launch = TRANSITIONS_PER_LAUNCH
    q0 = init_positions(rank, chains, DIM)
    s = gm.HMC(gm.RosenbrockND(DIM), q0, STEP_SIZE, N_LEAPFROG, seed=42, ctx=ctx, chain_offset=rank * chains)
    # warm-up with pooled dual averaging over all ranks (NCCL all-reduce of the acceptance statistic every transition)
    s.set_adaptation("pooled", 0.8)
    s.run_device(0, 100)
    plan = launches_for(steps, per_launch)
    wplan = launches_for(max(warmup, 3), per_launch)
    # the first warm-up launch is as long as the longest timed one, so the library-owned [chains, n, dim] sample buffer
    # reaches its final size before the timed region
    if max(plan) > max(wplan):
        wplan = [max(plan)] + wplan
    for n in wplan:
        s.run_device(n, 0)
    step_size = s.counters().step_size
    ms = env.timed(lambda: [s.run_device(n, 0) for n in plan])
    clk = env.clocks_now()
    unit_per_step = chains * N_LEAPFROG
    value = unit_per_step * steps * world / (ms * 1e-3)
    tfl = value / world * FLOP_PER_GRAD_EVAL / 1e12
    peak = env.fp32_peak()
    hbm = chains * BYTES_PER_STEP_PER_CHAIN * steps / (ms * 1e-3) / 1e9
    roof = {"bound": "fp32", "kernel": "hmc_run_kernel<float,25,RosenbrockND>", "achieved": tfl, "peak": peak, "unit": "TFLOP/s",
            "frac": tfl / peak if peak else None,
            "peak_source": "FFMA micro-benchmark in this run (gmcmc_measure_fp32_peak); nominal 148*128*2*%.3f GHz = %.1f"
                           % (env.pk["sm_max_mhz"] / 1e3, 148 * 128 * 2 * env.pk["sm_max_mhz"] / 1e6),
            "algorithmic_flop_per_unit": FLOP_PER_GRAD_EVAL,
            "hbm": {"achieved": hbm, "peak": env.pk["hbm_gbs"], "unit": "GB/s", "frac": hbm / env.pk["hbm_gbs"],
                    "algorithmic_bytes_per_unit": BYTES_PER_STEP_PER_CHAIN / N_LEAPFROG,
                    "note": "state is register-resident for all L steps; HBM carries only the sample write-out, so this "
                            "kernel is FP32-pipe bound, not HBM bound (north_star: FP32-pipe utilisation)"}}
    roof.update(traffic_fields("hmc_rosenbrock", chains))
    e2e_calls = max(8, min(20, steps // E2E_TRANSITIONS))     # >= 8 calls (~70 ms): PCIe rates wobble by several % between short bursts
    e2e, extra = e2e_leg(env, s, chains, DIM, np.float32, E2E_TRANSITIONS, e2e_calls,
                         units_per_call=unit_per_step * E2E_TRANSITIONS, variants=variants)
    e2e["unit"] = "grad-evals/s"
    for v in extra.values():
        v["unit"] = "grad-evals/s"

    # ---- min-ESS/sec (BASELINE metric, second half): n_ess draws per chain, ESS / R-hat reduced on the device
    ess = None
    if want_ess:
        n_ess = 500
        s.reserve(n_ess)
        holder = {}
        sample_ms = env.timed(lambda: holder.setdefault("p", s.run_device(n_ess, 0)))
        st = L.RunStatsC()
        for _ in range(2):       # the first call pays one-time kernel loading; the second is timed
            ctx.synchronize()
            t_s0 = time.perf_counter()
            L.check(lib.gmcmc_run_stats_from(ctx._h, C.c_void_p(holder["p"]), C.c_size_t(chains), C.c_size_t(n_ess), C.c_size_t(DIM),
                                             L.F32, 1, C.byref(st)))
            ctx.synchronize()
            t_s1 = time.perf_counter()
        stats_ms = env.max_over_ranks((t_s1 - t_s0) * 1e3)
        converged = st.rhat_std.max < 1.01
        ess = {"min_ess": st.ess.min, "median_ess": st.ess.median,
               "min_ess_per_sec": st.ess.min / (sample_ms * 1e-3) if converged else None,
               "draws_per_chain": n_ess, "chains_total": chains * world, "sampling_ms": sample_ms,
               "split_rhat_max": st.rhat_std.max, "converged": bool(converged),
               "device_stats_ms": stats_ms, "stats_read_gbs": chains * n_ess * DIM * 4 / (stats_ms * 1e-3) / 1e9,
               "note": "ESS per stats.rs:523-573 over ALL ranks' chains, reduced on the device (K4 + NCCL A2/A3); seconds = "
                       "sampling time of these draws (stats excluded), SURVEY 8(d).  min_ess_per_sec is null unless split "
                       "R-hat < 1.01: an ESS estimate of unconverged chains is not a throughput"}
    cpu = None
    if rank == 0 and world == 1 and want_cpu:
        r, th, sample = cpu_hmc_rate(10.0)
        cpu = {"value": r, "unit": "grad-evals/s", "cores": th, "kind": "port", "sample": sample}
    c = s.counters()
    out = {"metric": "leapfrog_grad_evals_per_sec", "value": value, "unit": "grad-evals/s", "ms_per_step": ms / steps,
           "dtype": "f32", "steps": steps,
           "config": {"workload": WORKLOAD_NAMES["hmc_rosenbrock"] % (DIM, chains, N_LEAPFROG), "chains_per_gpu": chains,
                      "transitions_per_launch": max(plan), "launch_plan": plan, "step_size": step_size,
                      "accept_rate": c.accept_rate,
                      "l2": "256 MB L2 flush before the timed region; chain state is register-resident within a launch and "
                            "each launch streams %.0f MB of samples (> 126 MB L2), so nothing is reused from L2 between "
                            "launches; ~4 ms of all-SM FP32 work (gmcmc_ctx_warm_fp32) is queued, untimed, ahead of the first event so the "
                            "region does not start from the idle state the barrier leaves" % (chains * BYTES_PER_STEP_PER_CHAIN * max(plan) / 1e6)},
           "e2e": e2e, "gpu_launches": len(plan), "roofline": roof, "cpu_baseline": cpu, "clocks": clk, "ess": ess}
    out.update(extra)
    s.close()
    return out


def leg_mh(env, steps, warmup, chains, want_cpu=True):
    gm, ctx = env.gm, env.ctx
    rank, world = env.rank, env.world
    per_launch = 1000
    x0 = np.random.default_rng(100 + rank).standard_normal((chains, 2))
    tgt = gm.Gaussian2D([0.0, 0.0], [[1.0, 0.0], [0.0, 1.0]])
    s = gm.MetropolisHastings(tgt, gm.IsotropicGaussian(1.0), x0, ctx=ctx, chain_offset=rank * chains).seed(42)
    plan = launches_for(steps, per_launch)
    wplan = launches_for(max(warmup, 3), per_launch)
    if max(plan) > max(wplan):
        wplan = [max(plan)] + wplan
    for n in wplan:
        s.run_device(n, 0)
    ms = env.timed(lambda: [s.run_device(n, 0) for n in plan])
    clk = env.clocks_now()
    value = chains * steps * world / (ms * 1e-3)
    ach = chains * 16 * steps / (ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": "mh_run2_kernel<double, Gaussian2D>", "achieved": ach, "peak": env.pk["hbm_gbs"], "unit": "GB/s",
            "frac": ach / env.pk["hbm_gbs"], "peak_source": env.pk["source"], "algorithmic_bytes_per_unit": 16,
            "note": "co-bound by instruction dispatch (Philox4x32-10 LOP3 / IMAD.WIDE, 13 FP64 and the selects sit on half-rate "
                    "pipes; one Philox block feeds two transitions); a write-only stream of the same 256-byte pieces reaches "
                    "5.3 TB/s (tools/microbench_write.cu)"}
    roof.update(traffic_fields("mh_gauss2d", chains))
    e2e, _ = e2e_leg(env, s, chains, 2, np.float64, 16, 4, units_per_call=chains * 16)
    e2e["unit"] = "chain-steps/s"
    cpu = None
    if rank == 0 and world == 1 and want_cpu:
        r, th, sample = cpu_mh_rate(10.0)
        cpu = {"value": r, "unit": "chain-steps/s", "cores": th, "kind": "port", "sample": sample}
    c = s.counters()
    out = {"metric": "mh_chain_steps_per_sec", "value": value, "unit": "chain-steps/s", "ms_per_step": ms / steps, "dtype": "f64",
           "steps": steps,
           "config": {"workload": WORKLOAD_NAMES["mh_gauss2d"] % chains, "chains_per_gpu": chains, "transitions_per_launch": per_launch,
                      "accept_rate": c.accept_rate,
                      "l2": "256 MB L2 flush before the timed region; every launch streams 16.8 GB of samples (> L2)"},
           "e2e": e2e, "gpu_launches": len(plan), "roofline": roof, "cpu_baseline": cpu, "clocks": clk}
    s.close()
    return out


def leg_dense(env, steps, warmup, chains, dd=1000, want_cpu=True):
    gm, ctx = env.gm, env.ctx
    rank, world = env.rank, env.world
    tgt = dense_target(gm, dd)
    q0 = np.random.default_rng(200 + rank).standard_normal((chains, dd)).astype(np.float32)
    s = gm.HMC(tgt, q0, 0.05, N_LEAPFROG, seed=42, ctx=ctx, chain_offset=rank * chains)
    per_launch = 2
    plan = launches_for(steps, per_launch)
    s.reserve(max(plan))
    for n in launches_for(max(warmup, 3), per_launch):
        s.run_device(n, 0)
    c0 = s.counters()
    ms = env.timed(lambda: [s.run_device(n, 0) for n in plan])
    clk = env.clocks_now()
    c1 = s.counters()
    value = chains * N_LEAPFROG * steps * world / (ms * 1e-3)
    flop = 2.0 * dd * dd
    tfl = value / world * flop / 1e12
    peak = env.pk.get("bf16_tflops_sustained", env.pk["bf16_tflops"])
    roof = {"bound": "tensor", "kernel": "dense_gemm_kick_kernel (tcgen05.mma kind::f16, 128x256x16, FP16 x 3 split)", "achieved": tfl,
            "peak": peak, "unit": "TFLOP/s", "frac": tfl / peak, "frac_of_tf32_peak": tfl / (peak / 2.0),
            "peak_source": "measured sustained 16-bit dense peak (%s)" % env.pk["source"],
            "algorithmic_flop_per_unit": flop, "executed_tensor_tflops": 3.0 * tfl * (1.0 + 1.0 / N_LEAPFROG),
            "note": "error-compensated FP16 split of both operands (hi.hi + lo.hi + hi.lo, FP32 accumulation in TMEM): the tensor "
                    "pipe executes 3x the algorithmic flops, and L+1 GEMMs per transition are credited as L"}
    roof.update(traffic_fields("hmc_dense", chains, dd == 1000))
    e2e, _ = e2e_leg(env, s, chains, dd, np.float32, 2, 3, units_per_call=chains * N_LEAPFROG * 2)
    e2e["unit"] = "grad-evals/s"
    cpu = None
    if rank == 0 and world == 1 and want_cpu:
        r, th, sample = cpu_dense_rate(10.0, dd)
        cpu = {"value": r, "unit": "grad-evals/s", "cores": th, "kind": "port", "sample": sample}
    out = {"metric": "leapfrog_grad_evals_per_sec", "value": value, "unit": "grad-evals/s", "ms_per_step": ms / steps,
           "dtype": "f32 (FP16 x 3 split tensor-core gradient, FP32 accumulation)", "steps": steps,
           "config": {"workload": WORKLOAD_NAMES["hmc_dense"] % (dd, chains, N_LEAPFROG), "chains_per_gpu": chains,
                      "transitions_per_launch": per_launch, "accept_rate": (c1.accepts - c0.accepts) / max(1, c1.transitions - c0.transitions),
                      "l2": "256 MB L2 flush before the timed region; p and delta (2 x 262 MB) exceed L2"},
           "e2e": e2e, "gpu_launches": int((c1.launches) * len(plan)), "roofline": roof, "cpu_baseline": cpu, "clocks": clk}
    s.close()
    return out


def leg_nuts(env, steps, warmup, chains, want_cpu=True):
    gm, ctx = env.gm, env.ctx
    rank, world = env.rank, env.world
    per_launch = 200          # tree sizes are heavy-tailed: long launches let the chain queue balance them
    w, mu = mixture_params()
    tgt = gm.GaussianMixture(w, mu, 1.0)
    q0 = np.random.default_rng(300 + rank).standard_normal((chains, DIM)).astype(np.float32)
    s = gm.NUTS(tgt, q0, 0.8, seed=42, ctx=ctx, chain_offset=rank * chains, max_depth=10)
    s.run_device(1, 100)              # warm-up: per-chain dual averaging (generic_nuts.rs:882-924)
    plan = launches_for(steps, per_launch)
    s.reserve(max(plan) + 1)
    for n in launches_for(max(warmup, 3), per_launch):
        s.run_device(n + 1, 0)
    c0 = s.counters()
    ms = env.timed(lambda: [s.run_device(n + 1, 0) for n in plan])
    clk = env.clocks_now()
    c1 = s.counters()
    units_local = c1.grad_evals - c0.grad_evals
    value = env.sum_over_ranks(units_local) / (ms * 1e-3)
    flop = (5 * 4 + 6) * DIM      # (5K + 6) d flop per leapfrog
    tfl = value / world * flop / 1e12
    peak = env.fp32_peak()
    roof = {"bound": "fp32", "kernel": "nuts_run_kernel<float, 8, TagMixture, padded, no mass, 16 lanes per chain>", "achieved": tfl, "peak": peak, "unit": "TFLOP/s",
            "frac": tfl / peak if peak else None, "peak_source": "FFMA micro-benchmark in this run", "algorithmic_flop_per_unit": flop,
            "mean_leapfrogs_per_transition": units_local / max(1, c1.transitions - c0.transitions),
            "note": "divergence-limited: chains of a warp build trees of different sizes (warp-level masking)"}
    roof.update(traffic_fields("nuts_mixture", chains))
    e2e, _ = e2e_leg(env, s, chains, DIM, np.float32, 4, 3, counter_units=True)
    e2e["unit"] = "grad-evals/s"
    # ---- min-ESS/sec (BASELINE metric, second half) on config 5 as stated: a fresh sampler, 200 warm-up + 200 collected
    # transitions inside the events, ESS / R-hat of the 200 draws reduced on the device over all ranks' chains
    L, lib = env.L, env.lib
    s2 = gm.NUTS(tgt, q0, 0.8, seed=43, ctx=ctx, chain_offset=rank * chains, max_depth=10)
    n_ess = 200
    s2.reserve(n_ess)
    holder = {}
    c_a = s2.counters()
    run_ms = env.timed(lambda: holder.setdefault("p", s2.run_device(n_ess, 200)))
    c_b = s2.counters()
    st = L.RunStatsC()
    for _ in range(2):
        ctx.synchronize()
        t_s0 = time.perf_counter()
        L.check(lib.gmcmc_run_stats_from(ctx._h, C.c_void_p(holder["p"]), C.c_size_t(chains), C.c_size_t(n_ess), C.c_size_t(DIM),
                                         L.F32, 1, C.byref(st)))
        ctx.synchronize()
        t_s1 = time.perf_counter()
    stats_ms = env.max_over_ranks((t_s1 - t_s0) * 1e3)
    converged = st.rhat_std.max < 1.01
    ess = {"min_ess": st.ess.min, "median_ess": st.ess.median,
           "min_ess_per_sec": st.ess.min / (run_ms * 1e-3) if converged else None,
           "draws_per_chain": n_ess, "warmup_transitions": 200, "chains_total": chains * world, "run_ms": run_ms,
           "leapfrogs_per_sec_incl_warmup": env.sum_over_ranks(c_b.grad_evals - c_a.grad_evals) / (run_ms * 1e-3),
           "split_rhat_max": st.rhat_std.max, "converged": bool(converged), "device_stats_ms": stats_ms,
           "note": "config 5 as stated (200 warm-up + 200 collected transitions, both inside the timed region); ESS per "
                   "stats.rs:523-573 over all ranks' chains; seconds = the whole run, burn-in included, stats excluded (SURVEY 8d)"}
    s2.close()
    cpu = None
    if rank == 0 and world == 1 and want_cpu:
        r, th, sample = cpu_nuts_rate(10.0)
        cpu = {"value": r, "unit": "grad-evals/s", "cores": th, "kind": "port", "sample": sample}
    out = {"metric": "leapfrog_grad_evals_per_sec", "value": value, "unit": "grad-evals/s", "ms_per_step": ms / steps, "dtype": "f32",
           "steps": steps, "ess": ess,
           "config": {"workload": WORKLOAD_NAMES["nuts_mixture"] % (DIM, chains), "chains_per_gpu": chains,
                      "transitions_per_launch": per_launch, "step_size": c1.step_size,
                      "accept_rate": (c1.accepts - c0.accepts) / max(1, c1.transitions - c0.transitions),
                      "divergences": c1.divergences - c0.divergences,
                      "accept_rate_meaning": "fraction of transitions that accepted a subtree proposal (the chain moved)",
                      "l2": "256 MB L2 flush before the timed region"},
           "e2e": e2e, "gpu_launches": len(plan), "roofline": roof, "cpu_baseline": cpu, "clocks": clk}
    s.close()
    return out


def leg_cfg4_strong(env, total_chains=STRONG_CHAINS, n_collect=1000, n_discard=200):
    """BASELINE config 4 as stated: `total_chains` sharded over the ranks, pooled dual-averaging warm-up (collective A1
    every transition) + n_collect draws inside the timed region, device RunStats over all ranks' chains after it."""
    gm, L, lib, ctx = env.gm, env.L, env.lib, env.ctx
    rank, world = env.rank, env.world
    lo, hi = gm.shard_chains(total_chains, rank, world)
    chains = hi - lo
    rng = np.random.default_rng(9000 + rank)
    q0 = (1.0 + 0.1 * rng.standard_normal((chains, DIM))).astype(np.float32)
    s = gm.HMC(gm.RosenbrockND(DIM), q0, STEP_SIZE, N_LEAPFROG, seed=42, ctx=ctx, chain_offset=lo)
    s.set_adaptation("pooled", 0.8)
    s.reserve(n_collect)                        # the [chains, n, dim] buffer (105 GB at one GPU) is sized before the clock starts
    s.run_device(4, 4)                          # kernels loaded, NCCL warmed up (8 untimed transitions)
    s.set_positions(q0)
    s.set_step_size(STEP_SIZE)
    s.set_adaptation("pooled", 0.8)
    holder = {}

    def work(mark):
        s.run_device(0, n_discard)
        mark()
        holder["p"] = s.run_device(n_collect, 0)
    ms, (warm_ms, collect_ms) = env.timed(work, marks=1)
    clk = env.clocks_now()
    c = s.counters()
    value = total_chains * (n_collect + n_discard) * N_LEAPFROG / (ms * 1e-3)
    st = L.RunStatsC()
    stats_times = []
    for _ in range(2):       # the first call pays one-time costs (kernel loading, work-buffer allocation); the second is reported
        ctx.synchronize()
        t0 = time.perf_counter()
        L.check(lib.gmcmc_run_stats_from(ctx._h, C.c_void_p(holder["p"]), C.c_size_t(chains), C.c_size_t(n_collect), C.c_size_t(DIM),
                                         L.F32, 1, C.byref(st)))
        ctx.synchronize()
        stats_times.append(env.max_over_ranks((time.perf_counter() - t0) * 1e3))
    stats_ms = stats_times[-1]
    converged = st.rhat_std.max < 1.01
    tfl = value / world * FLOP_PER_GRAD_EVAL / 1e12
    peak = env.fp32_peak()
    out = {"metric": "leapfrog_grad_evals_per_sec", "value": value, "unit": "grad-evals/s", "scaling": "strong",
           "chains_total": total_chains, "chains_per_gpu": chains, "n_discard": n_discard, "n_collect": n_collect,
           "ms_total": ms, "warmup_ms": warm_ms, "collect_ms": collect_ms,
           "warmup_ms_per_transition": warm_ms / n_discard, "collect_ms_per_transition": collect_ms / n_collect,
           "warmup_cost_ratio": (warm_ms / n_discard) / (collect_ms / n_collect),
           "step_size": c.step_size, "roofline_frac_fp32": tfl / peak if peak else None,
           "collectives_in_timed_region": "A1: %d NCCL all-reduces of 2 doubles (one per warm-up transition, on a side stream, "
                                          "one transition lagged)" % (n_discard if world > 1 else 0),
           "device_stats_ms": stats_ms, "device_stats_first_call_ms": stats_times[0],
           "stats_read_gbs": chains * n_collect * DIM * 4 / (stats_ms * 1e-3) / 1e9,
           "min_ess": st.ess.min, "median_ess": st.ess.median, "split_rhat_max": st.rhat_std.max, "converged": bool(converged),
           "min_ess_per_sec": st.ess.min / (ms * 1e-3) if converged else None, "clocks": clk,
           "note": "value = total chains x (warm-up + collected) transitions x L / device time of the whole run (max over "
                   "ranks); the driver's scaling efficiency for config 4 as stated follows from this object's value over N"}
    s.close()
    return out


def leg_g_invariant(env):
    """A sharded run reproduces the unsharded chains bit for bit: Philox streams are keyed by the global chain index."""
    gm, gdist, ctx = env.gm, env.gdist, env.ctx
    rank, world = env.rank, env.world
    Ct, d, n = 4096, 20, 48
    q0 = (1.0 + 0.05 * np.random.default_rng(0).standard_normal((Ct, d))).astype(np.float32)
    parts = max(world, 2)
    if world > 1:
        lo, hi = gm.shard_chains(Ct, rank, world)
        s = gm.HMC(gm.RosenbrockND(d), q0[lo:hi], 0.01, 8, seed=42, ctx=ctx, chain_offset=lo)
        mine = gdist.chain_checksums(s.run(n, 5))
        s.close()
        sums = np.concatenate(gdist.all_gather_object(mine))
    else:
        pieces = []
        for r in range(parts):
            lo, hi = gm.shard_chains(Ct, r, parts)
            s = gm.HMC(gm.RosenbrockND(d), q0[lo:hi], 0.01, 8, seed=42, ctx=ctx, chain_offset=lo)
            pieces.append(gdist.chain_checksums(s.run(n, 5)))
            s.close()
        sums = np.concatenate(pieces)
    ok = None
    if rank == 0:
        solo = gm.Context(env.local) if world > 1 else ctx
        ref = gm.HMC(gm.RosenbrockND(d), q0, 0.01, 8, seed=42, ctx=solo)
        ok = bool(np.array_equal(sums, gdist.chain_checksums(ref.run(n, 5))))
        ref.close()
    return ok


def run_ours(args, rank, world, local):
    env = Env(rank, world, local)
    want_cpu = not args.no_cpu
    wl = args.workload
    line = None
    try:
        if wl in (None, "hmc_rosenbrock"):
            head = leg_hmc_rosenbrock(env, args.steps, args.warmup, args.chains or CHAINS_PER_GPU, want_ess=not args.no_ess,
                                      want_cpu=want_cpu)
        elif wl == "mh_gauss2d":
            head = leg_mh(env, args.steps, args.warmup, args.chains or 1048576, want_cpu)
        elif wl == "hmc_dense":
            head = leg_dense(env, args.steps, args.warmup, args.chains or 65536, args.dim or 1000, want_cpu)
        else:
            head = leg_nuts(env, args.steps, args.warmup, args.chains or 65536, want_cpu)
        line = {"metric": head.pop("metric"), "value": head.pop("value"), "unit": head.pop("unit"), "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": head.pop("ms_per_step"),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": head.pop("dtype"), "data": "synthetic"}
        head.pop("steps", None)
        line.update(head)
        if world > 1:
            line["rank_cores"] = len(env.cores) if env.cores else None
        if wl is None and not args.headline_only:
            legs = {}
            for name, fn in (("mh_gauss2d", lambda: leg_mh(env, 2000, 1000, 1048576, want_cpu)),
                             ("hmc_dense", lambda: leg_dense(env, 4, 3, 65536, 1000, want_cpu)),
                             ("nuts_mixture", lambda: leg_nuts(env, 200, 20, 65536, want_cpu))):
                try:
                    legs[name] = fn()
                    legs[name]["n_gpus"] = world
                    legs[name]["scaling"] = "weak"
                except Exception as e:   # a failed leg is reported, never silently dropped
                    legs[name] = {"error": repr(e)}
                env.barrier()
            line["workloads"] = legs
            try:
                line["cfg4_strong"] = leg_cfg4_strong(env)
            except Exception as e:
                line["cfg4_strong"] = {"error": repr(e)}
            env.barrier()
            try:
                line["g_invariant"] = leg_g_invariant(env)
            except Exception as e:
                line["g_invariant"] = None
                line["g_invariant_error"] = repr(e)
    finally:
        env.clocks.stop()
    if rank == 0 and line is not None:
        print(json.dumps(line), flush=True)
    if world > 1:
        env.dist.barrier()
        env.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4000)
    ap.add_argument("--warmup", type=int, default=400)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=["hmc_rosenbrock", "mh_gauss2d", "nuts_mixture", "hmc_dense"],
                    help="print this workload alone (default: the headline + every other config nested in one line)")
    ap.add_argument("--headline-only", action="store_true", help="default workload without the nested legs")
    ap.add_argument("--dim", type=int, default=0, help="dimension (hmc_dense only; default 1000)")
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default: the workload's)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-ess", action="store_true", help="skip the min-ESS/sec leg")
    args = ap.parse_args()
    rank, world, local = dist_env()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, world, local)


if __name__ == "__main__":
    main()
