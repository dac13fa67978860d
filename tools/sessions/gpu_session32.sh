#!/bin/bash
# K4: twiddle table + global-memory lock step (progress words): parity tests, wall time and DRAM reads per variant
set -u
out=gpurun_out; mkdir -p $out
timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -q -x -p no:cacheprovider -k "stat or rhat or ess" 2>&1 | tail -3
for cfg in "0 1 4" "1 1 4" "1 1 1" "1 1 16" "1 4 8" "1 8 64"; do
  set -- $cfg
  echo "== lockstep=$1 sync_every=$2 slack=$3"
  GMCMC_STATS_LOCKSTEP=$1 GMCMC_STATS_SYNC_EVERY=$2 GMCMC_STATS_SLACK=$3 timeout 200 python tools/stats_bench.py 2>&1 | grep -E "wall|K4"
done
for cfg in "0 1 4" "1 1 4" "1 4 8"; do
  set -- $cfg
  echo "== ncu lockstep=$1 sync_every=$2 slack=$3"
  GMCMC_STATS_LOCKSTEP=$1 GMCMC_STATS_SYNC_EVERY=$2 GMCMC_STATS_SLACK=$3 timeout 300 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,launch__grid_size,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:stats_accumulate_warp -c 1 python tools/stats_bench.py 2>&1 | grep -E "dram__bytes_read|gpu__time|hit_rate|grid|issue_active|inst_executed" | tail -6
done
for l in 1 0; do
GMCMC_STATS_LOCKSTEP=$l timeout 200 python tools/stats_bench.py 65536 1000 100 2>&1 | grep -E "K4"
GMCMC_STATS_LOCKSTEP=$l timeout 200 python tools/stats_bench.py 65536 200 100 2>&1 | grep -E "K4"
done
