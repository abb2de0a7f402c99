#!/bin/bash
set -u
OUT=gpurun_out/r02p; mkdir -p $OUT
run_bench() {
  name=$1; shift
  ( env $ENVV timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-checks "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err )
  python - "$OUT/bench_$name.json" "$name" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms_per_step %.1f e2e %.1f"%(d["ms_per_step"], d["e2e"]["ms_per_step"]), d["roofline"].get("train_step_timed"), d["engine"][:75])
except Exception as ex:
    print(sys.argv[2], "FAILED", ex)
PY
}
export DI_BENCH_PREDICTORS=0
: > $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_GROUPS=8 DEEPIMPUTE_B200_ADAM_TPC=2" run_bench g8_t2 >> $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_GROUPS=8 DEEPIMPUTE_B200_ADAM_TPC=3" run_bench g8_t3 >> $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_GROUPS=10 DEEPIMPUTE_B200_ADAM_TPC=3" run_bench g10_t3 >> $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_GROUPS=20 DEEPIMPUTE_B200_ADAM_TPC=2" run_bench g20_t2 >> $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_GROUPS=13 DEEPIMPUTE_B200_ADAM_TPC=3" run_bench g13_t3 >> $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_GROUPS=5 DEEPIMPUTE_B200_ADAM_TPC=4" run_bench g5_t4 >> $OUT/summary.txt
ENVV="A=1" run_bench c5 --workload c5 >> $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_LT=0 DEEPIMPUTE_B200_ADAM=pers" run_bench c3_shard2_conv_pers --emulate-shard 0/2 >> $OUT/summary.txt
cat $OUT/summary.txt
