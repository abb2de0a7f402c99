#!/bin/bash
set -u
OUT=gpurun_out/r02w; mkdir -p $OUT
for B in 256 64; do
  for X in 0 1; do
    echo "== batch $B exact_math $X"
    FA_BATCH=$B DEEPIMPUTE_B200_EXACT_MATH=$X timeout 600 python scripts/family_accuracy.py 2>&1 | grep -v "^Epoch" | cut -c1-130 | grep "^lt \|conv ts  \|fp32"
  done
done | tee $OUT/exact_math_accuracy.txt
export DI_BENCH_PREDICTORS=0
for X in 0 1; do
  DEEPIMPUTE_B200_EXACT_MATH=$X timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/bench_c3_x$X.json 2> $OUT/bench_c3_x$X.err
  python - $OUT/bench_c3_x$X.json $X <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d["roofline"]; pc=d.get("parity_check") or {}
print("exact", sys.argv[2], "ms_per_step %.1f"%d["ms_per_step"], {n:v["ms"] for n,v in r["kernels"].items()}, pc.get("max_rel"), pc.get("max_rel_weights"))
PY
done
DEEPIMPUTE_B200_EXACT_MATH=1 timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --emulate-shard 0/8 > $OUT/bench_shard8_x1.json 2> $OUT/bench_shard8_x1.err
python -c "
import json;d=json.loads(open('$OUT/bench_shard8_x1.json').read().strip().splitlines()[-1]);print('shard8 exact ms_per_step %.1f'%d['ms_per_step'], d.get('parity_check',{}).get('max_rel'), d.get('parity_check',{}).get('max_rel_weights'))"
