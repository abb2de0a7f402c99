#!/bin/bash
# One gpurun call: GPU parity tests, smoke, both bench arms, ncu launch list of one c3 epoch.
# usage (repo root on the GPU box): bash scripts/gpu_check.sh [tag]
tag=${1:-check}
out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $out/gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)" >> $out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > $out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1; echo "smoke exit $?" >> $out/smoke.txt
timeout 900 python bench.py > $out/bench_c3.json 2> $out/bench_c3.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err
bash scripts/gpu_launchlist.sh $tag c3 tf32x3 > $out/launch_summary.txt 2>&1
tail -n 14 $out/pytest_gpu.txt; cat $out/smoke.txt; cat $out/bench_c3.json $out/bench_ref.json; tail -n 3 $out/*.err; cat $out/launch_summary.txt
