#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
for cfg in "1 1 1" "1 4 2" "1 4 3"; do
  set -- $cfg
  echo "== cluster=$1 sync_every=$2 waves=$3"
  GMCMC_STATS_DEBUG=1 GMCMC_STATS_CLUSTER=$1 GMCMC_STATS_SYNC_EVERY=$2 GMCMC_STATS_WAVES=$3 timeout 200 python tools/stats_bench.py 2>&1 | grep -E "wall|K4|stats\]" | sort | uniq | head -5
done
for cfg in "0 1 1" "1 4 1"; do
  set -- $cfg
  echo "== ncu cluster=$1 sync_every=$2"
  GMCMC_STATS_CLUSTER=$1 GMCMC_STATS_SYNC_EVERY=$2 timeout 300 ncu --metrics launch__grid_size,launch__cluster_size,launch__occupancy_cluster_gpu_pct,launch__occupancy_cluster_max_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__ctas_launched.sum,smsp__warp_issue_stalled_barrier_per_warp_active.pct,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_membar_per_warp_active.pct,gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.avg --clock-control none -k regex:stats_accumulate_warp -c 1 python tools/stats_bench.py 2>&1 | grep -E "launch__|sm__|smsp__|gpu__" | tail -14
done
