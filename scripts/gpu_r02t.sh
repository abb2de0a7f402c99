#!/bin/bash
set -u
OUT=gpurun_out/r02t; mkdir -p $OUT
( timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q -x 2>&1 | tail -5 ) > $OUT/pytest_engine.txt; tail -3 $OUT/pytest_engine.txt
( DEEPIMPUTE_B200_LT=0 timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q -x 2>&1 | tail -5 ) > $OUT/pytest_engine_conv.txt; tail -3 $OUT/pytest_engine_conv.txt
run_bench() {
  name=$1; shift
  ( env $ENVV timeout 1200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err )
  python - "$OUT/bench_$name.json" "$name" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d["roofline"]["kernels"]; pc=d.get("parity_check") or {}
    print(sys.argv[2], "ms_per_step %.1f e2e %.1f"%(d["ms_per_step"], d["e2e"]["ms_per_step"]), {n:(v["ms"]) for n,v in k.items()}, pc.get("max_rel"), pc.get("max_rel_weights"), d["roofline"].get("train_step_timed"), d["engine"][:80])
except Exception as ex:
    print(sys.argv[2], "FAILED", ex)
PY
}
export DI_BENCH_PREDICTORS=0
: > $OUT/summary.txt
ENVV="DI_BENCH_ORACLE_C5=1" run_bench c5 --workload c5 >> $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_ADAM=ring" run_bench c5_ring --workload c5 --no-checks >> $OUT/summary.txt
ENVV="A=1" run_bench c5_shard8 --workload c5 --emulate-shard 0/8 --no-checks >> $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_ADAM=ring" run_bench c5_shard8_ring --workload c5 --emulate-shard 0/8 --no-checks >> $OUT/summary.txt
cat $OUT/summary.txt
