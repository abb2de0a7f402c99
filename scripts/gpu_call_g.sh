#!/bin/bash
tag=${1:-s3g}
out=gpurun_out/$tag; mkdir -p $out
timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_multinet_gpu.py -m gpu -q > $out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.txt
DI_BENCH_PREDICTORS=0 timeout 600 python bench.py --steps 2 --warmup 1 --epochs 5 --no-cpu-baseline > $out/ab_hoist.json 2> $out/ab_hoist.err
DEEPIMPUTE_B200_TRACE=1 DEEPIMPUTE_B200_DEEP=1 timeout 120 python scripts/trace_step.py step tf32x3 > $out/trace_x3.txt 2>&1
tail -n 3 $out/pytest_gpu.txt; python -c "
import json; d=json.load(open('$out/ab_hoist.json')); print('ms/step(5 epochs+predict)', d['ms_per_step'])"; grep "trace " $out/trace_x3.txt | tail -4
