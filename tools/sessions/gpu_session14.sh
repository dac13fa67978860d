#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout -s KILL 600 python -m pytest tests/test_gpu_gibbs.py -q -x -p no:cacheprovider > $out/r2_pytest_gibbs.txt 2>&1; tail -15 $out/r2_pytest_gibbs.txt
timeout -s KILL 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $out/r2_pytest_s14.txt 2>&1; tail -3 $out/r2_pytest_s14.txt
for v in "" _nuts5 _nuts6; do
  GMCMC_LIB=general_mcmc_b200/libgmcmc$v.so timeout 400 python bench.py --workload nuts_mixture --steps 200 --warmup 20 --no-cpu > $out/nuts_variant$v.json 2> $out/nuts_variant$v.err
  python - "$out/nuts_variant$v.json" "$v" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("nuts variant", sys.argv[2] or "base", d["value"], d["ms_per_step"])
except Exception as e:
    print("unreadable", sys.argv[1], e)
PY
done
