#!/bin/bash
# K1 launch head/tail: flat 16-byte state load / store for unpadded rows
set -u
out=gpurun_out; mkdir -p $out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_custom_target.py -q -x -p no:cacheprovider 2>&1 | tail -2
timeout 300 python tools/k1_launch_scan.py 2>&1 | tail -12
timeout 200 python tools/k1_rate.py
