#!/bin/bash
# K1 gradient cache (fast mode, RosenbrockND): parity tests, A/B of the steady-state rate and the per-launch scan
set -u
out=gpurun_out; mkdir -p $out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -q -x -p no:cacheprovider -k "hmc or rosen or continuation or shard" 2>&1 | tail -3
for i in 1 2; do
  timeout 200 python tools/k1_rate.py
  GMCMC_LIB=$PWD/general_mcmc_b200/libgmcmc_nogc.so timeout 200 python tools/k1_rate.py
done
timeout 300 python tools/k1_launch_scan.py 2>&1 | tail -14
