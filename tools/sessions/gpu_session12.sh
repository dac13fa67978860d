#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout -s KILL 900 python -m pytest tests/test_gpu_dense_tc.py -m gpu -q -s -p no:cacheprovider > $out/r2_pytest_s12.txt 2>&1; grep -E "GPU err|bench grid|cov scale|passed|failed|FAILED|^E  " $out/r2_pytest_s12.txt | cut -c1-200 | head -30
timeout 300 python bench.py --workload hmc_dense --steps 6 --warmup 3 --no-cpu > $out/r2_bench_dense3.json 2> $out/r2_bench_dense3.err; python - $out/r2_bench_dense3.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "dense", d["value"], d["roofline"]["frac"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("accept_rate"))
except Exception as e:
    print(sys.argv[1], "unreadable", e); print(open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
timeout 300 ncu --set full --import-source on --clock-control none -k regex:dense_gemm_kick -s 40 -c 1 -o $out/r2_full_dense_c -f python bench.py --workload hmc_dense --steps 4 --warmup 3 --no-cpu > $out/ncu_dense_full.log 2>&1; tail -2 $out/ncu_dense_full.log
