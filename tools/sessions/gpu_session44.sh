#!/bin/bash
# final records of round 2 on one GPU: bench lines (default flags, the driver's flags, reference arm), launch list, K1 capture
set -u
out=gpurun_out; mkdir -p $out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $out/r2_pytest_gpu_full.txt 2>&1; tail -1 $out/r2_pytest_gpu_full.txt
timeout 300 python __graft_entry__.py --smoke > $out/r2_smoke.txt 2>&1; tail -1 $out/r2_smoke.txt | cut -c1-200
timeout 900 python bench.py > $out/r2_bench_default.json 2> $out/r2_bench_default.err
timeout 900 python bench.py --steps 20 --warmup 5 > $out/r2_bench_default_s20.json 2> $out/r2_bench_default_s20.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $out/r2_bench_reference.json 2> $out/r2_bench_reference.err
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
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/r2_launches_default_bench.csv python bench.py --steps 20 --warmup 5 --no-cpu > $out/ncu_launches.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:hmc_run_kernel -s 45 -c 1 -o $out/r2_full_hmc -f python bench.py --headline-only --steps 3000 --warmup 100 --no-cpu --no-ess > $out/ncu_hmc.log 2>&1
ls -la $out/r2_full_hmc.ncu-rep $out/r2_launches_default_bench.csv
