#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout -s KILL 900 python -m pytest tests/test_gpu_dense_tc.py tests/test_gpu_parity.py -m gpu -q -s -p no:cacheprovider > $out/r2_pytest_s11.txt 2>&1; grep -E "GPU err|bench grid|cov scale|passed|failed|FAILED|^E  " $out/r2_pytest_s11.txt | cut -c1-200 | head -30
for w in 1 2 3 4 0; do
GMCMC_DENSE_WAVES=$w timeout 300 python bench.py --workload hmc_dense --steps 6 --warmup 3 --no-cpu > $out/r2_bench_dense_w$w.json 2> $out/r2_bench_dense_w$w.err; python - $out/r2_bench_dense_w$w.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "dense", d["value"], d["roofline"]["frac"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("accept_rate"))
except Exception as e:
    print(sys.argv[1], "unreadable", e); print(open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
done
