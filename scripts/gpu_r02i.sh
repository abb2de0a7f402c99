#!/bin/bash
# N = 1 knob sweep on c3 (graph mode): sub-network groups, ADAM kernel, ring depth, release point.
set -u
OUT=gpurun_out/r02i; mkdir -p $OUT
run_bench() {
  name=$1; shift
  ( env $ENVV timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-checks "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err )
  python - "$OUT/bench_$name.json" "$name" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms_per_step %.1f e2e %.1f"%(d["ms_per_step"], d["e2e"]["ms_per_step"]), d["roofline"].get("train_step_timed"), d.get("engine"))
except Exception as ex:
    print(sys.argv[2], "FAILED", ex)
PY
}
export DI_BENCH_PREDICTORS=0
: > $OUT/summary.txt
ENVV="A=1" run_bench base >> $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_GROUPS=8" run_bench g8 >> $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_GROUPS=24" run_bench g24 >> $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_GROUPS=40" run_bench g40 >> $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_ADAM=ring" run_bench ring >> $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_DEEP=0" run_bench deep0 >> $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_PDL_LEAD=2" run_bench lead2 >> $OUT/summary.txt
cat $OUT/summary.txt
