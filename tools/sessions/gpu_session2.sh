#!/bin/bash
# round 2, session 2: re-run of the NUTS / MH parity tests with their diagnostics, NUTS bench + quick ncu counters, K1 scan
set -u
out=gpurun_out; mkdir -p $out
timeout -s KILL 600 python -m pytest tests/test_gpu_nuts.py tests/test_gpu_parity.py tests/test_gpu_dense_tc.py -m gpu -q -s -p no:cacheprovider > $out/r2_pytest_s2.txt 2>&1; grep -E "cfg5|mh2 |cov scale|GPU err|passed|failed|FAILED" $out/r2_pytest_s2.txt | head -40
timeout 300 python bench.py --workload nuts_mixture --steps 200 --warmup 20 --no-cpu > $out/r2_bench_nuts.json 2> $out/r2_bench_nuts.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_nuts.json").read().strip().splitlines()[-1])
    print("nuts", d["value"], d["roofline"]["frac"], d["roofline"].get("mean_leapfrogs_per_transition"), d["config"].get("accept_rate"), d["e2e"]["value"])
except Exception as e:
    print("nuts bench unreadable", e); print(open("gpurun_out/r2_bench_nuts.err").read()[-2000:])
PY
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread --clock-control none -k regex:nuts_run_kernel -s 1 -c 1 --csv --log-file $out/r2_nuts_counters.csv python bench.py --workload nuts_mixture --steps 200 --warmup 20 --no-cpu > $out/ncu_nuts_q.log 2>&1; tail -8 $out/r2_nuts_counters.csv | cut -c1-300
timeout 300 python tools/k1_scan.py > $out/r2_k1_scan2.txt 2>&1; cat $out/r2_k1_scan2.txt
