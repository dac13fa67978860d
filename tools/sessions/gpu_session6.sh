#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 200 python tools/k1_launch_scan.py > $out/r2_k1_launch_scan2.txt 2>&1; cat $out/r2_k1_launch_scan2.txt
timeout 300 python tools/k1_scan.py > $out/r2_k1_scan3.txt 2>&1; cat $out/r2_k1_scan3.txt
timeout -s KILL 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $out/r2_pytest_s6.txt 2>&1; tail -3 $out/r2_pytest_s6.txt
