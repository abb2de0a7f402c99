#!/bin/bash
# One gpurun call: building-block probe, GPU parity tests, smoke, a bench line and an ncu launch list.
# usage (from the repo root on the GPU box): bash scripts/gpu_round.sh [tag]
tag=${1:-r1}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $out/gpu.txt 2>&1
for a in 0 1; do for b in 0 1; do
  timeout 60 deepimpute_b200/csrc/umma_probe $a $b >> $out/umma_probe.txt 2>&1; echo "exit $?" >> $out/umma_probe.txt
done; done
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1; echo "smoke exit $?" >> $out/smoke.txt
timeout 900 python bench.py --workload c2 --steps 2 --warmup 1 > $out/bench_c2.json 2> $out/bench_c2.err
timeout 1500 python bench.py > $out/bench_c3.json 2> $out/bench_c3.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_c2.csv \
  python bench.py --workload c2 --steps 1 --warmup 0 --epochs 1 --no-cpu-baseline > $out/ncu_bench.log 2>&1
tail -3 $out/umma_probe.txt $out/pytest_gpu.txt $out/smoke.txt
cat $out/bench_c2.json $out/bench_c3.json
tail -5 $out/bench_c3.err
