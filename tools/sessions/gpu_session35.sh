#!/bin/bash
# ncu --set full of ONE single-transition K1 launch (where does the ~40 us per-launch fixed cost go)
set -u
out=gpurun_out; mkdir -p $out
timeout 300 python tools/k1_one.py > /dev/null 2>&1 || echo "k1_one failed"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:hmc_run_kernel -s 4 -c 1 -f -o $out/r2_k1_single_after python tools/k1_one.py > $out/ncu_k1_single.log 2>&1; tail -3 $out/ncu_k1_single.log
