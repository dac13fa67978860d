#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout -s KILL 600 python -m pytest tests/test_gpu_nuts.py tests/test_gpu_parity.py -m gpu -q -s -p no:cacheprovider > $out/r2_pytest_s5.txt 2>&1; grep -E "passed|failed|FAILED|^E  " $out/r2_pytest_s5.txt | cut -c1-250 | head -20
for cfg in "8 16" "13 8" "25 4"; do set -- $cfg
GMCMC_NUTS_EPL=$1 GMCMC_NUTS_LPC=$2 timeout 300 python bench.py --workload nuts_mixture --steps 200 --warmup 20 --no-cpu > $out/r2_bench_nuts_$1_$2.json 2> $out/r2_bench_nuts_$1_$2.err; python - $out/r2_bench_nuts_$1_$2.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "nuts", d["value"], d["roofline"]["frac"], d["roofline"].get("mean_leapfrogs_per_transition"), d["e2e"]["value"])
except Exception as e:
    print(sys.argv[1], "nuts bench unreadable", e)
PY
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:hmc_run_kernel -s 3 -c 1 -o $out/r2_full_k1_one -f python tools/k1_one.py > $out/ncu_k1_one.log 2>&1; tail -2 $out/ncu_k1_one.log
