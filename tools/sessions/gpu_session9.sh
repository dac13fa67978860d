#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout -s KILL 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > $out/r2_pytest_s9.txt 2>&1; grep -E "dense mass|passed|failed|FAILED|^E  " $out/r2_pytest_s9.txt | cut -c1-300 | head -60
