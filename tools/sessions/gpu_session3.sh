#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout -s KILL 600 python -m pytest tests -m gpu -q -s -p no:cacheprovider > $out/r2_pytest_s3.txt 2>&1; grep -E "cfg5|passed|failed|FAILED|^E  " $out/r2_pytest_s3.txt | cut -c1-250 | head -40
for v in "" _minb4; do
GMCMC_LIB=$PWD/general_mcmc_b200/libgmcmc$v.so timeout 300 python bench.py --workload nuts_mixture --steps 200 --warmup 20 --no-cpu > $out/r2_bench_nuts$v.json 2> $out/r2_bench_nuts$v.err; python - $out/r2_bench_nuts$v.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "nuts", d["value"], d["roofline"]["frac"], d["roofline"].get("mean_leapfrogs_per_transition"), d["config"].get("accept_rate"), d["e2e"]["value"])
except Exception as e:
    print("nuts bench unreadable", e)
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none -k regex:nuts_run_kernel -s 1 -c 1 --csv --log-file $out/r2_nuts_counters3.csv python bench.py --workload nuts_mixture --steps 200 --warmup 20 --no-cpu > $out/ncu_nuts_q.log 2>&1; tail -5 $out/r2_nuts_counters3.csv | cut -d, -f13-15
timeout 400 python tools/k1_window_scan.py > $out/r2_k1_window_scan.txt 2>&1; cat $out/r2_k1_window_scan.txt
