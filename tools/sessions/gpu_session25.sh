#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout -s KILL 600 python -m pytest tests/test_gpu_dense_tc.py -q -x -p no:cacheprovider > $out/r2_pytest_s25.txt 2>&1; tail -3 $out/r2_pytest_s25.txt
for i in 1 2; do
timeout 400 python bench.py --workload hmc_dense --steps 6 --warmup 3 --no-cpu > $out/dense_s25.json 2> $out/dense_s25.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/dense_s25.json").read().strip().splitlines()[-1])
    print("dense", d["value"], d["ms_per_step"], d["clocks"])
except Exception as e:
    print("unreadable", e)
PY
done
