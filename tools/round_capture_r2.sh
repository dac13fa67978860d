#!/bin/bash
# tools/round_capture_r2.sh — one GPU session that refreshes the round-2 evidence under profiles/: the GPU test suite, smoke,
# the default bench line and the reference arm (never under a profiler), the launch list of the default bench and one
# `ncu --set full` capture per kernel.  Run from the repo root:  gpurun --timeout 2400 -- 'bash tools/round_capture_r2.sh'
set -u
out=gpurun_out; mkdir -p $out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $out/r2_pytest_gpu_full.txt 2>&1; tail -3 $out/r2_pytest_gpu_full.txt
timeout 300 python __graft_entry__.py --smoke > $out/r2_smoke.txt 2>&1; tail -2 $out/r2_smoke.txt | cut -c1-300
timeout 900 python bench.py > $out/r2_bench_default.json 2> $out/r2_bench_default.err; tail -c 300 $out/r2_bench_default.err
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
timeout 200 python tools/stats_bench.py > $out/r2_stats_bench.txt 2>&1; tail -4 $out/r2_stats_bench.txt
timeout 300 python tools/mh_dims.py > $out/r2_mh_dims.txt 2>&1; cat $out/r2_mh_dims.txt
# launch list of the default driver-style run (cold-cache, serialised times: the kernels' SHARES are what it is for)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/r2_launches_default_bench.csv python bench.py --steps 20 --warmup 5 --no-cpu > $out/ncu_launches.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:hmc_run_kernel -s 45 -c 1 -o $out/r2_full_hmc -f python bench.py --headline-only --steps 3000 --warmup 100 --no-cpu --no-ess > $out/ncu_hmc.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:stats_accumulate_warp -s 1 -c 1 -o $out/r2_full_stats_warp -f python tools/stats_bench.py > $out/ncu_stats.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:mh_run2 -s 2 -c 1 -o $out/r2_full_mh2 -f python bench.py --workload mh_gauss2d --steps 2000 --warmup 1000 --no-cpu > $out/ncu_mh2.log 2>&1
ls -la $out/*.ncu-rep | tail -6
