#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 200 python tools/k1_launch_scan.py > $out/r2_k1_launch_scan.txt 2>&1; cat $out/r2_k1_launch_scan.txt
timeout 400 ncu --set full --import-source on --clock-control none -k regex:nuts_run_kernel -s 1 -c 1 -o $out/r2_full_nuts_a -f python bench.py --workload nuts_mixture --steps 200 --warmup 20 --no-cpu > $out/ncu_nuts_full.log 2>&1; tail -3 $out/ncu_nuts_full.log
ls -la $out/r2_full_nuts_a.ncu-rep
