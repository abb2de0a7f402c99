#!/bin/bash
set -u
OUT=gpurun_out/r02e; mkdir -p $OUT
( timeout 300 python scripts/tc_vs_fp32_epoch.py 2>&1 | grep -v "^Epoch" ) > $OUT/tc_vs_fp32_epoch.txt
cat $OUT/tc_vs_fp32_epoch.txt
