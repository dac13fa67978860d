#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 400 ncu --set full --import-source on --clock-control none -k regex:nuts_run_kernel -s 1 -c 1 -o $out/r2_full_nuts_c -f python bench.py --workload nuts_mixture --steps 40 --warmup 20 --no-cpu > $out/ncu_nuts.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:hmc_run_kernel -s 45 -c 1 -o $out/r2_full_hmc -f python bench.py --headline-only --steps 3000 --warmup 100 --no-cpu --no-ess > $out/ncu_hmc.log 2>&1
ls -la $out/r2_full_nuts_c.ncu-rep $out/r2_full_hmc.ncu-rep
