#!/bin/bash
# round 2, session 1: GPU test suite (with the diagnostics the new parity tests print), smoke, the full default bench line,
# K1 scan.  Run on the GPU box from the repo root.
set -u
out=gpurun_out; mkdir -p $out
timeout -s KILL 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider 2>&1 | tail -120 > $out/r2_pytest_gpu.txt; tail -40 $out/r2_pytest_gpu.txt
timeout 300 python __graft_entry__.py --smoke > $out/r2_smoke.txt 2>&1; tail -3 $out/r2_smoke.txt
timeout 600 python bench.py > $out/r2_bench_default.json 2> $out/r2_bench_default.err; tail -c 1500 $out/r2_bench_default.err
timeout 300 python tools/k1_scan.py > $out/r2_k1_scan.txt 2>&1; cat $out/r2_k1_scan.txt
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_default.json").read().strip().splitlines()[-1])
    print("headline", d["value"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], d.get("e2e_stats_only", {}).get("value"), d.get("e2e_device", {}).get("value"))
    print("ess", d.get("ess"))
    for k, v in d.get("workloads", {}).items():
        print(k, v.get("value"), (v.get("roofline") or {}).get("frac"), (v.get("e2e") or {}).get("value"), (v.get("cpu_baseline") or {}).get("value"), v.get("error"))
    print("cfg4_strong", d.get("cfg4_strong"))
    print("g_invariant", d.get("g_invariant"), d.get("g_invariant_error"))
except Exception as e:
    print("bench line unreadable:", e)
PY
