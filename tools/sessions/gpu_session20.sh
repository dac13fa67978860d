#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout -s KILL 900 python -m pytest tests/test_gpu_stats.py tests/test_gpu_edges.py -q -x -p no:cacheprovider > $out/r2_pytest_s20.txt 2>&1; tail -3 $out/r2_pytest_s20.txt
timeout 200 python tools/stats_bench.py > $out/r2_stats_bench.txt 2>&1; tail -4 $out/r2_stats_bench.txt
timeout 200 python tools/stats_bench.py 65536 1000 100 > $out/r2_stats_bench_1000.txt 2>&1; tail -4 $out/r2_stats_bench_1000.txt
