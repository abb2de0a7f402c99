#!/bin/bash
# One gpurun call: building-block probe, GPU parity tests, smoke, bench lines and an ncu launch list.
# usage (from the repo root on the GPU box): bash scripts/gpu_round.sh [tag] [bench workloads...]
tag=${1:-r1}; shift
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $out/gpu.txt 2>&1
for a in 0 1; do for b in 0 1; do
  timeout 60 deepimpute_b200/csrc/umma_probe $a $b >> $out/umma_probe.txt 2>&1; echo "exit $?" >> $out/umma_probe.txt
done; done
timeout 1500 python -m pytest tests -m gpu -q -s > $out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1; echo "smoke exit $?" >> $out/smoke.txt
for spec in "$@"; do   # spec = workload:math
  w=${spec%%:*}; m=${spec##*:}
  timeout 1500 python bench.py --workload $w --math $m > $out/bench_${w}_${m}.json 2> $out/bench_${w}_${m}.err
done
cat $out/umma_probe.txt; tail -n 40 $out/pytest_gpu.txt; cat $out/smoke.txt
for f in $out/bench_*.json; do cat $f; done
for f in $out/bench_*.err; do tail -n 5 $f; done
