#!/bin/bash
# eight GPUs of one box: the default bench under torchrun (weak-scaling headline, every workload, cfg4_strong with the NCCL
# warm-up collective inside the timed region, g_invariant), then the cross-rank acceptance script
set -u
out=gpurun_out; mkdir -p $out
N=${1:-8}
nvidia-smi -L | wc -l; nproc; nvidia-smi topo -m 2>/dev/null | head -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > $out/r2_bench_${N}gpu.json 2> $out/r2_bench_${N}gpu.err; tail -c 400 $out/r2_bench_${N}gpu.err
python - $N <<'PY'
import json, sys
N = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_%sgpu.json" % N).read().strip().splitlines() if l.startswith("{")][-1])
    print("headline", d["value"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], d.get("e2e_stats_only", {}).get("value"), d.get("e2e_device", {}).get("value"), "cores/rank", d.get("rank_cores"))
    print("ess", {k: (d.get("ess") or {}).get(k) for k in ("min_ess", "split_rhat_max", "device_stats_ms")})
    for k, v in d.get("workloads", {}).items():
        print(k, v.get("value"), (v.get("roofline") or {}).get("frac"), (v.get("e2e") or {}).get("value"), v.get("error"))
    c = d.get("cfg4_strong", {})
    print("cfg4_strong", {k: c.get(k) for k in ("value", "chains_per_gpu", "ms_total", "warmup_ms", "collect_ms", "warmup_cost_ratio", "step_size", "device_stats_ms", "split_rhat_max", "roofline_frac_fp32", "error")})
    print("g_invariant", d.get("g_invariant"), d.get("g_invariant_error"))
except Exception as e:
    print("bench line unreadable:", e)
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > $out/r2_multigpu_check_${N}gpu.txt 2>&1; grep -v "^W\|^\[W\|^\*\*\|OMP_NUM" $out/r2_multigpu_check_${N}gpu.txt | tail -6
