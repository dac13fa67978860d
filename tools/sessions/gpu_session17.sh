#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $out/r2_pytest_s17.txt 2>&1; tail -3 $out/r2_pytest_s17.txt
timeout 300 python tools/k1_launch_scan.py > $out/r2_k1_launch_scan_epsconst.txt 2>&1; grep '"chains": 65536' $out/r2_k1_launch_scan_epsconst.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $out/r2_bench_default_s20.json 2> $out/r2_bench_default_s20.err
timeout 900 python bench.py > $out/r2_bench_default.json 2> $out/r2_bench_default.err
python - <<'PY'
import json
for f in ("r2_bench_default", "r2_bench_default_s20"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "headline", d["value"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], "ess", (d.get("ess") or {}).get("device_stats_ms"))
        for k, v in d.get("workloads", {}).items():
            print("  ", k, v.get("value"), (v.get("roofline") or {}).get("frac"), v.get("error"))
        c = d.get("cfg4_strong", {})
        print("   cfg4_strong", c.get("value"), c.get("warmup_cost_ratio"), "g_invariant", d.get("g_invariant"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
