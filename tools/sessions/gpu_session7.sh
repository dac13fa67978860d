#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout -s KILL 900 python -m pytest tests/test_gpu_dense_tc.py -m gpu -q -s -p no:cacheprovider > $out/r2_pytest_s7.txt 2>&1; grep -E "GPU err|bench grid|cov scale|passed|failed|^E  " $out/r2_pytest_s7.txt | cut -c1-200
for v in "" _epi12 _epi16; do
GMCMC_LIB=$PWD/general_mcmc_b200/libgmcmc$v.so timeout 300 python bench.py --workload hmc_dense --steps 6 --warmup 3 --no-cpu > $out/r2_bench_dense$v.json 2> $out/r2_bench_dense$v.err; python - $out/r2_bench_dense$v.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "dense", d["value"], d["roofline"]["frac"], d["ms_per_step"], d["e2e"]["value"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
for v in ""; do GMCMC_LIB=$PWD/general_mcmc_b200/libgmcmc$v.so timeout 120 python tools/k1_rate.py; done
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:dense_gemm_kick -s 40 -c 3 --csv --log-file $out/r2_dense_counters.csv python bench.py --workload hmc_dense --steps 4 --warmup 3 --no-cpu > $out/ncu_dense_q.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_dense_counters.csv')) if len(r)>10 and r[0].isdigit()]
for i in sorted(set(r[0] for r in rows)):
    print(i, {r[-3]:r[-1] for r in rows if r[0]==i})
PY
