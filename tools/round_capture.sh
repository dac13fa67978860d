#!/bin/bash
# tools/round_capture.sh — one GPU session that refreshes everything under profiles/: the GPU test suite, the four
# bench lines (never under a profiler), the launch list of the default bench and one `ncu --set full` capture per kernel.
# Run on the GPU box from the repo root:  gpurun --timeout 1500 -- 'bash tools/round_capture.sh'
set -u
out=gpurun_out; mkdir -p $out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $out/r1_pytest_gpu.txt; cat $out/r1_pytest_gpu.txt
timeout 400 python bench.py > $out/bench_r1_hmc.json 2> $out/bench_r1_hmc.err
timeout 300 python bench.py --workload mh_gauss2d --steps 4000 --warmup 1000 > $out/bench_r1_mh.json 2> $out/bench_r1_mh.err
timeout 400 python bench.py --workload nuts_mixture --steps 200 --warmup 20 > $out/bench_r1_nuts.json 2> $out/bench_r1_nuts.err
timeout 400 python bench.py --workload hmc_dense --steps 6 --warmup 3 > $out/bench_r1_dense.json 2> $out/bench_r1_dense.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_r1_reference.json 2> $out/bench_r1_reference.err
for f in hmc mh nuts dense reference; do python - $out/bench_r1_$f.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d.get("value"), (d.get("roofline") or {}).get("frac"), (d.get("e2e") or {}).get("value"), (d.get("ess") or {}).get("device_stats_ms"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/r1_launches_hmc_bench.csv python bench.py --steps 300 --warmup 100 --no-cpu > $out/ncu_launches.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:stats_accumulate_warp -s 1 -c 1 -o $out/r1_full_stats_warp -f python tools/stats_bench.py > $out/ncu_stats.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:mh_run2 -s 2 -c 1 -o $out/r1_full_mh2 -f python bench.py --workload mh_gauss2d --steps 2000 --warmup 1000 --no-cpu > $out/ncu_mh2.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:nuts_run_kernel -s 1 -c 1 -o $out/r1_full_nuts -f python bench.py --workload nuts_mixture --steps 200 --warmup 20 --no-cpu > $out/ncu_nuts.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:dense_gemm_kick -s 40 -c 1 -o $out/r1_full_dense -f python bench.py --workload hmc_dense --steps 4 --warmup 3 --no-cpu > $out/ncu_dense.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:hmc_run_kernel -s 101 -c 1 -o $out/r1_full_hmc -f python bench.py --steps 300 --warmup 100 --no-cpu --no-ess > $out/ncu_hmc.log 2>&1
ls -la $out/*.ncu-rep | tail -6
