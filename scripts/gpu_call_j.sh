#!/bin/bash
# Programmatic dependent launch along the step chain: correctness and effect at N = 1.
tag=${1:-s3j}
out=gpurun_out/$tag; mkdir -p $out
DEEPIMPUTE_B200_PDL=1 timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_multinet_gpu.py -m gpu -q > $out/pytest_pdl.txt 2>&1; echo "pytest exit $?" >> $out/pytest_pdl.txt
run() { name=$1; shift
  env DI_BENCH_PREDICTORS=0 "$@" timeout 600 python bench.py --steps 2 --warmup 1 --epochs 5 --no-cpu-baseline > $out/ab_$name.json 2> $out/ab_$name.err
  python -c "
import json; d=json.load(open('$out/ab_$name.json')); print('%-10s ms/step(5 epochs+predict) %.2f' % ('$name', d['ms_per_step']))" 2>&1 | tail -1
}
run pdl1 DEEPIMPUTE_B200_PDL=1 > $out/ab.txt
run pdl0 DEEPIMPUTE_B200_PDL=0 >> $out/ab.txt
run pdl1_g40 DEEPIMPUTE_B200_PDL=1 DEEPIMPUTE_B200_GROUPS=40 >> $out/ab.txt
grep -E "passed|failed|exit" $out/pytest_pdl.txt; grep -E "^FAILED|^E  " $out/pytest_pdl.txt | head; cat $out/ab.txt; tail -n 3 $out/ab_pdl1.err
