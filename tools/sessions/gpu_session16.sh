#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
for v in "" _epsc "" _epsc; do
  GMCMC_LIB=general_mcmc_b200/libgmcmc$v.so timeout 300 python tools/k1_launch_scan.py 2>&1 | grep '"chains": 65536' | grep -E 'launch": (32|128),' | sed "s/^/variant[$v] /"
done
./tools/microbench_fp32 > $out/r2_microbench_fp32.txt 2>&1; cat $out/r2_microbench_fp32.txt
