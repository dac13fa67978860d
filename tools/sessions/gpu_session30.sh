#!/bin/bash
# K4 cluster lock-step experiment: parity tests, wall time and DRAM reads per variant
set -u
out=gpurun_out; mkdir -p $out
timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -q -x -p no:cacheprovider -k "stat or rhat or ess" 2>&1 | tail -3
for cfg in "0 1" "1 1" "1 2" "1 4" "1 16"; do
  set -- $cfg
  echo "== cluster=$1 sync_every=$2"
  GMCMC_STATS_CLUSTER=$1 GMCMC_STATS_SYNC_EVERY=$2 timeout 200 python tools/stats_bench.py 2>&1 | grep -E "wall|K4"
done
for cfg in "0 1" "1 1" "1 4"; do
  set -- $cfg
  echo "== ncu cluster=$1 sync_every=$2"
  GMCMC_STATS_CLUSTER=$1 GMCMC_STATS_SYNC_EVERY=$2 timeout 300 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:stats_accumulate_warp -c 2 python tools/stats_bench.py 2>&1 | grep -E "dram__bytes_read|gpu__time|hit_rate|grid|Cluster" | tail -6
done
GMCMC_STATS_CLUSTER=1 timeout 200 python tools/stats_bench.py 65536 1000 100 2>&1 | grep -E "K4"
GMCMC_STATS_CLUSTER=0 timeout 200 python tools/stats_bench.py 65536 1000 100 2>&1 | grep -E "K4"
GMCMC_STATS_CLUSTER=1 timeout 200 python tools/stats_bench.py 65536 200 100 2>&1 | grep -E "K4"
GMCMC_STATS_CLUSTER=0 timeout 200 python tools/stats_bench.py 65536 200 100 2>&1 | grep -E "K4"
