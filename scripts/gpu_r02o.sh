#!/bin/bash
# Persistent ADAM kernel: parity (converter family forced on the small shapes), then A/B on c3.
set -u
OUT=gpurun_out/r02o; mkdir -p $OUT
( timeout 600 python scripts/family_accuracy.py 2>&1 | grep -v "^Epoch" | tail -30 ) > $OUT/family_accuracy.txt
cut -c1-200 $OUT/family_accuracy.txt | grep "conv ts \|^lt \|fp32"
( DEEPIMPUTE_B200_LT=0 timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_multinet_gpu.py tests/test_full_size_gpu.py -m gpu -q -x 2>&1 | tail -6 ) > $OUT/pytest_conv_family.txt
tail -3 $OUT/pytest_conv_family.txt
( timeout 900 python -m pytest tests/test_benchmarked_parity_gpu.py -m gpu -q 2>&1 | tail -6 ) > $OUT/pytest_bench_parity.txt
tail -3 $OUT/pytest_bench_parity.txt
run_bench() {
  name=$1; shift
  ( env $ENVV timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err )
  python - "$OUT/bench_$name.json" "$name" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d["roofline"]["kernels"]; pc=d.get("parity_check") or {}
    print(sys.argv[2], "ms_per_step %.1f e2e %.1f"%(d["ms_per_step"], d["e2e"]["ms_per_step"]), {n:(v["ms"]) for n,v in k.items()}, pc.get("max_rel"), pc.get("max_rel_weights"), d["roofline"].get("train_step_timed"), d["engine"][:60])
except Exception as ex:
    print(sys.argv[2], "FAILED", ex)
PY
}
export DI_BENCH_PREDICTORS=0
: > $OUT/summary.txt
ENVV="A=1" run_bench c3_pers2 >> $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_ADAM_TPC=3" run_bench c3_pers3 --no-checks >> $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_ADAM_TPC=4" run_bench c3_pers4 --no-checks >> $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_ADAM_TPC=1" run_bench c3_pers1 --no-checks >> $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_ADAM=big" run_bench c3_big --no-checks >> $OUT/summary.txt
cat $OUT/summary.txt
