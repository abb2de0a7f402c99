#!/bin/bash
# ncu evidence for profiles/: (1) launch list with device time per launch, (2) one --set full capture of a kernel.
# usage: bash scripts/gpu_profile.sh <tag> <workload> <kernel-regex>
tag=${1:-prof}; wl=${2:-c3}; kre=${3:-tc_adam_kernel}
out=gpurun_out/$tag
mkdir -p $out
cmd="python bench.py --workload $wl --steps 1 --warmup 0 --epochs 1 --no-cpu-baseline"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tc_|gather_kernel|simt_' -s 100 -c 800 \
    --csv --log-file $out/launches_$wl.csv $cmd > $out/launches_$wl.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:$kre -s 20 -c 2 \
    -o $out/full_${wl}_${kre} $cmd > $out/full_$wl.log 2>&1
ls -la $out
tail -n 3 $out/launches_$wl.log $out/full_$wl.log
