#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 400 ncu --set full --import-source on --clock-control none -k regex:nuts_run_kernel -s 1 -c 1 -o $out/r2_full_nuts_d -f python bench.py --workload nuts_mixture --steps 40 --warmup 20 --no-cpu > $out/ncu_nuts.log 2>&1
ls -la $out/r2_full_nuts_d.ncu-rep
