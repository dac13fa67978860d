#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 300 python __graft_entry__.py --smoke > $out/r2_smoke.txt 2>&1; tail -2 $out/r2_smoke.txt | cut -c1-500
timeout -s KILL 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $out/r2_pytest_gpu_full.txt 2>&1; tail -3 $out/r2_pytest_gpu_full.txt
