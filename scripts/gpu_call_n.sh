#!/bin/bash
tag=${1:-s3n}
out=gpurun_out/$tag; mkdir -p $out
run() { name=$1; shift
  env DI_BENCH_PREDICTORS=0 "$@" timeout 600 python bench.py --steps 2 --warmup 1 --epochs 5 --no-cpu-baseline > $out/ab_$name.json 2> $out/ab_$name.err
  python -c "
import json; d=json.load(open('$out/ab_$name.json')); print('%-10s ms/step(5 epochs+predict) %.2f adam %.1f us' % ('$name', d['ms_per_step'], d['roofline']['kernels']['adam']['ms']*1e3))" 2>&1 | tail -1
}
run spec DEEPIMPUTE_B200_ADAM_GENERIC=0 > $out/ab.txt
run generic DEEPIMPUTE_B200_ADAM_GENERIC=1 >> $out/ab.txt
run spec2 DEEPIMPUTE_B200_ADAM_GENERIC=0 >> $out/ab.txt
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q > $out/pytest.txt 2>&1; echo "pytest exit $?" >> $out/pytest.txt
cat $out/ab.txt; grep -E "passed|failed|exit" $out/pytest.txt
