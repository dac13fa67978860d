#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout -s KILL 900 python -m pytest tests/test_gpu_nuts.py tests/test_gpu_parity.py -q -x -p no:cacheprovider > $out/r2_pytest_s26.txt 2>&1; tail -3 $out/r2_pytest_s26.txt
grep "cfg5 mixture" $out/r2_pytest_s26.txt | head
for i in 1 2; do
timeout 400 python bench.py --workload nuts_mixture --steps 200 --warmup 20 --no-cpu > $out/nuts_s26.json 2> $out/nuts_s26.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/nuts_s26.json").read().strip().splitlines()[-1])
    print("nuts", d["value"], d["ms_per_step"], d["ess"]["min_ess_per_sec"], d["ess"]["split_rhat_max"])
except Exception as e:
    print("unreadable", e)
PY
done
