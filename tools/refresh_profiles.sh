#!/bin/bash
# copies the records of the last tools/round_capture_r2.sh session from gpurun_out/ into profiles/ and regenerates the
# ncu text summaries and the launch-list summary
set -e
cd "$(dirname "$0")/.."
cp gpurun_out/r2_launches_default_bench.csv profiles/
cp gpurun_out/r2_bench_default.json profiles/r2_bench_default_1gpu.json
cp gpurun_out/r2_bench_default_s20.json profiles/r2_bench_default_1gpu_steps20.json
cp gpurun_out/r2_bench_reference.json profiles/r2_bench_reference_arm.json
cp gpurun_out/r2_pytest_gpu_full.txt profiles/r2_pytest_gpu.txt
cp gpurun_out/r2_smoke.txt profiles/r2_smoke.txt
cp gpurun_out/r2_mh_dims.txt profiles/r2_mh_generic_dims.txt
cp gpurun_out/r2_stats_bench.txt profiles/r2_stats_bench.txt
for k in hmc:r2_hmc_run_kernel_full mh2:r2_mh_run2_kernel_full stats_warp:r2_stats_accumulate_warp_full; do
  bash profiles/_summarize.sh gpurun_out/r2_full_${k%%:*}.ncu-rep profiles/${k##*:}.txt 14
done
python - <<'PY' > profiles/r2_launches_default_bench_summary.txt
import csv, collections, re
rows = list(csv.reader(l for l in open("profiles/r2_launches_default_bench.csv") if not l.startswith("==")))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    if len(r) <= vi: continue
    name = re.sub(r"\(.*", "", r[ki])[:80]
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
    tot[name] += v; cnt[name] += 1
T = sum(tot.values())
print("ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 20 --warmup 5 --no-cpu (default line: headline + cfg2/3/5 legs + cfg4_strong + g_invariant)")
print("cold-cache, serialised per-launch times: the kernels' SHARES are what this list is for.  %d launches, %.1f ms of GPU time" % (sum(cnt.values()), T / 1e3))
for k, v in tot.most_common(40):
    print("%-82s n=%5d  %9.2f ms  %5.1f%%" % (k, cnt[k], v / 1e3, 100 * v / T))
PY
head -8 profiles/r2_launches_default_bench_summary.txt
