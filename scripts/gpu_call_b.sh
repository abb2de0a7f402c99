#!/bin/bash
# GPU tests incl. the fused ends, bench with the postprocess block, pipeline traces, full ncu captures for profiles/.
tag=${1:-s3b}
out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > $out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.txt
timeout 900 python bench.py > $out/bench_c3.json 2> $out/bench_c3.err
for m in tf32x3 tf32; do
  DEEPIMPUTE_B200_TRACE=1 timeout 120 python scripts/trace_step.py step $m > $out/trace_$m.txt 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_adam_kernel -s 40 -c 1 -o $out/full_c3_adam \
    python bench.py --steps 1 --warmup 0 --epochs 1 --no-cpu-baseline > $out/full_c3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'impute_kernel|counts_to_norm' -c 3 -o $out/full_post_impute \
    python scripts/trace_step.py impute > $out/full_impute.log 2>&1
tail -n 12 $out/pytest_gpu.txt; cat $out/bench_c3.json; tail -n 3 $out/bench_c3.err; tail -n 60 $out/trace_tf32x3.txt; tail -n 4 $out/full_c3.log $out/full_impute.log; ls -la $out
