#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 400 python bench.py --workload nuts_mixture --steps 200 --warmup 20 --no-cpu > $out/nuts_ess.json 2> $out/nuts_ess.err; tail -c 300 $out/nuts_ess.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/nuts_ess.json").read().strip().splitlines()[-1])
print(d["value"], d["ess"])
PY
timeout 900 python bench.py --steps 20 --warmup 5 > $out/r2_bench_default_s20.json 2> $out/r2_bench_default_s20.err; tail -c 300 $out/r2_bench_default_s20.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_default_s20.json").read().strip().splitlines()[-1])
print(d["value"], d["workloads"]["nuts_mixture"]["ess"])
PY
