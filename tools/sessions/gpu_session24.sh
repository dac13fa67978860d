#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $out/r2_pytest_s24.txt 2>&1; tail -3 $out/r2_pytest_s24.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $out/r2_bench_default_s20.json 2> $out/r2_bench_default_s20.err; tail -c 300 $out/r2_bench_default_s20.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_default_s20.json").read().strip().splitlines()[-1])
print("headline", d["value"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e_stats_only"], d["e2e_device"])
for k, v in d["workloads"].items(): print(k, v.get("value"), v.get("error"))
print(d["cfg4_strong"]["value"], d["g_invariant"])
PY
