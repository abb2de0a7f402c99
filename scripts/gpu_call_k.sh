#!/bin/bash
tag=${1:-s3k}
out=gpurun_out/$tag; mkdir -p $out
run() { name=$1; shift
  env DI_BENCH_PREDICTORS=0 "$@" timeout 600 python bench.py --steps 2 --warmup 1 --epochs 5 --no-cpu-baseline > $out/ab_$name.json 2> $out/ab_$name.err
  python -c "
import json; d=json.load(open('$out/ab_$name.json')); print('%-10s ms/step(5 epochs+predict) %.2f' % ('$name', d['ms_per_step']))" 2>&1 | tail -1
}
run pdl2 DEEPIMPUTE_B200_PDL=2 > $out/ab.txt
run pdl1 DEEPIMPUTE_B200_PDL=1 >> $out/ab.txt
run pdl2_g8 DEEPIMPUTE_B200_PDL=2 DEEPIMPUTE_B200_GROUPS=8 >> $out/ab.txt
DEEPIMPUTE_B200_PDL=2 timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q -k "epochs" > $out/pytest_pdl2.txt 2>&1; echo "pytest exit $?" >> $out/pytest_pdl2.txt
cat $out/ab.txt; grep -E "passed|failed|exit" $out/pytest_pdl2.txt
