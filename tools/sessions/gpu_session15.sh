#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $out/r2_pytest_s15.txt 2>&1; tail -3 $out/r2_pytest_s15.txt
timeout 600 python bench.py --headline-only --no-cpu > $out/k1_packed_default.json 2> $out/k1_packed_default.err
timeout 600 python bench.py --headline-only --no-cpu --steps 20 --warmup 5 > $out/k1_packed_s20.json 2> $out/k1_packed_s20.err
python - <<'PY'
import json
for f in ("k1_packed_default", "k1_packed_s20"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], d["roofline"]["frac"], d["ms_per_step"], "peak", d["roofline"]["peak"], "clocks", d["clocks"])
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 120 python tools/k1_launch_scan.py > $out/k1_launch_scan_packed.txt 2>&1; tail -12 $out/k1_launch_scan_packed.txt
