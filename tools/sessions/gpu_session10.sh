#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout -s KILL 900 python -m pytest tests/test_gpu_nuts.py tests/test_gpu_custom_target.py -m gpu -q -s -p no:cacheprovider > $out/r2_pytest_s10.txt 2>&1; grep -E "cfg5|dense mass f|passed|failed|FAILED|^E  " $out/r2_pytest_s10.txt | cut -c1-250 | head -40
timeout 300 python bench.py --workload nuts_mixture --steps 200 --warmup 20 --no-cpu > $out/r2_bench_nuts3.json 2> $out/r2_bench_nuts3.err; python - $out/r2_bench_nuts3.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "nuts", d["value"], d["roofline"]["frac"], d["roofline"].get("mean_leapfrogs_per_transition"), d["e2e"]["value"])
except Exception as e:
    print(sys.argv[1], "nuts bench unreadable", e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
timeout 400 ncu --set full --import-source on --clock-control none -k regex:nuts_run_kernel -s 1 -c 1 -o $out/r2_full_nuts_b -f python bench.py --workload nuts_mixture --steps 200 --warmup 20 --no-cpu > $out/ncu_nuts_full.log 2>&1; tail -2 $out/ncu_nuts_full.log
