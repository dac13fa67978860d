#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $out/r2_pytest_s13.txt 2>&1; tail -3 $out/r2_pytest_s13.txt
timeout 300 python __graft_entry__.py --smoke > $out/r2_smoke.txt 2>&1; tail -2 $out/r2_smoke.txt | cut -c1-400
timeout 900 python bench.py --steps 20 --warmup 5 > $out/r2_bench_default_s20.json 2> $out/r2_bench_default_s20.err; tail -c 600 $out/r2_bench_default_s20.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $out/r2_bench_reference.json 2> $out/r2_bench_reference.err; tail -c 300 $out/r2_bench_reference.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_default_s20.json").read().strip().splitlines()[-1])
    print("headline", d["value"], d["roofline"]["frac"], d["ms_per_step"], "e2e", d["e2e"]["value"], d.get("e2e_stats_only", {}).get("value"), d.get("e2e_device", {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    print("ess", {k: (d.get("ess") or {}).get(k) for k in ("min_ess", "split_rhat_max", "device_stats_ms", "stats_read_gbs")})
    for k, v in d.get("workloads", {}).items():
        print(k, v.get("value"), (v.get("roofline") or {}).get("frac"), (v.get("e2e") or {}).get("value"), (v.get("cpu_baseline") or {}).get("value"), v.get("error"))
    c = d.get("cfg4_strong", {})
    print("cfg4_strong", {k: c.get(k) for k in ("value", "ms_total", "warmup_ms", "collect_ms", "warmup_cost_ratio", "step_size", "device_stats_ms", "split_rhat_max", "roofline_frac_fp32", "error")})
    print("g_invariant", d.get("g_invariant"), d.get("g_invariant_error"))
    r = json.loads(open("gpurun_out/r2_bench_reference.json").read().strip().splitlines()[-1])
    print("reference", r["value"], r["cpu_baseline"])
except Exception as e:
    print("bench line unreadable:", e)
PY
