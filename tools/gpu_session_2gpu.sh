#!/bin/bash
# two GPUs of one box: the cross-rank acceptance script (G-invariance, RunStats over ranks, pooled adaptation) and the
# bench under torchrun (weak-scaling headline + cfg4_strong with the NCCL warm-up collective inside the timed region)
set -u
out=gpurun_out; mkdir -p $out
nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -s -p no:cacheprovider > $out/r2_pytest_multi.txt 2>&1; tail -5 $out/r2_pytest_multi.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > $out/r2_multigpu_check_2gpu.txt 2>&1; grep -v "^W\|^\[W" $out/r2_multigpu_check_2gpu.txt | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2000 --warmup 200 > $out/r2_bench_2gpu.json 2> $out/r2_bench_2gpu.err; tail -c 600 $out/r2_bench_2gpu.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_2gpu.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("headline", d["value"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], d.get("e2e_stats_only", {}).get("value"), d.get("e2e_device", {}).get("value"))
    for k, v in d.get("workloads", {}).items():
        print(k, v.get("value"), (v.get("roofline") or {}).get("frac"), (v.get("e2e") or {}).get("value"), v.get("error"))
    c = d.get("cfg4_strong", {})
    print("cfg4_strong", {k: c.get(k) for k in ("value", "ms_total", "warmup_ms", "collect_ms", "warmup_cost_ratio", "step_size", "device_stats_ms", "split_rhat_max", "error")})
    print("g_invariant", d.get("g_invariant"), d.get("g_invariant_error"), "ess", (d.get("ess") or {}).get("device_stats_ms"))
except Exception as e:
    print("bench line unreadable:", e)
PY
