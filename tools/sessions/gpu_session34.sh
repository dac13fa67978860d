#!/bin/bash
# K1 launch-head state load: all loads in flight (default) against 8 per round (libgmcmc_un8.so)
set -u
out=gpurun_out; mkdir -p $out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -q -x -p no:cacheprovider -k "hmc or rosen or continuation or shard" 2>&1 | tail -2
for i in 1 2; do
echo "== default (UN=32)"; timeout 300 python tools/k1_launch_scan.py 2>&1 | grep 65536
echo "== UN=8"; GMCMC_LIB=$PWD/general_mcmc_b200/libgmcmc_un8.so timeout 300 python tools/k1_launch_scan.py 2>&1 | grep 65536
done
