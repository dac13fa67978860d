#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
for v in _pipe ""; do
GMCMC_LIB=general_mcmc_b200/libgmcmc$v.so timeout -s KILL 300 python -m pytest tests/test_gpu_dense_tc.py -q -x -p no:cacheprovider 2>&1 | tail -2
GMCMC_LIB=general_mcmc_b200/libgmcmc$v.so timeout 300 python bench.py --workload hmc_dense --steps 6 --warmup 3 --no-cpu > $out/dense_s29.json 2> $out/dense_s29.err
python - "$v" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/dense_s29.json").read().strip().splitlines()[-1])
    print("dense", sys.argv[1] or "base", d["value"], d["ms_per_step"], d["clocks"])
except Exception as e:
    print("unreadable", e)
PY
done
