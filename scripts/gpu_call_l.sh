#!/bin/bash
tag=${1:-s3l}
out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_multinet_gpu.py -m gpu -q > $out/pytest.txt 2>&1; echo "pytest exit $?" >> $out/pytest.txt
run() { name=$1; shift
  env DI_BENCH_PREDICTORS=0 "$@" timeout 600 python bench.py --steps 2 --warmup 1 --epochs 5 --no-cpu-baseline > $out/ab_$name.json 2> $out/ab_$name.err
  python -c "
import json; d=json.load(open('$out/ab_$name.json')); print('%-10s ms/step(5 epochs+predict) %.2f' % ('$name', d['ms_per_step']))" 2>&1 | tail -1
}
run pre1 DEEPIMPUTE_B200_PDL_PREFETCH=1 > $out/ab.txt
run pre0 DEEPIMPUTE_B200_PDL_PREFETCH=0 >> $out/ab.txt
run pre1_lead2 DEEPIMPUTE_B200_PDL_PREFETCH=1 DEEPIMPUTE_B200_PDL_LEAD=2 >> $out/ab.txt
run pre1_lead4 DEEPIMPUTE_B200_PDL_PREFETCH=1 DEEPIMPUTE_B200_PDL_LEAD=4 >> $out/ab.txt
cat $out/ab.txt; grep -E "passed|failed|exit" $out/pytest.txt; grep -E "^FAILED|^E  " $out/pytest.txt | head
