#!/bin/bash
set -u
for v in "" _u1 _u4 "" _u1 _u4; do
  GMCMC_LIB=general_mcmc_b200/libgmcmc$v.so timeout 300 python tools/k1_launch_scan.py 2>&1 | grep '"chains": 65536' | grep -E 'launch": (128),' | sed "s/^/variant[$v] /"
done
