#!/bin/bash
# Converter teams A/B (N = 1 kernel family).
set -u
OUT=gpurun_out/r02m; mkdir -p $OUT
( DEEPIMPUTE_B200_LT=0 timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_multinet_gpu.py tests/test_full_size_gpu.py -m gpu -q -x 2>&1 | tail -6 ) > $OUT/pytest_conv_family.txt
tail -4 $OUT/pytest_conv_family.txt
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
    print(sys.argv[2], "ms_per_step %.1f e2e %.1f"%(d["ms_per_step"], d["e2e"]["ms_per_step"]), {n:(v["ms"]) for n,v in k.items()}, d["roofline"].get("predict",{}).get("ms"), pc.get("max_rel"), pc.get("max_rel_weights"), d["roofline"].get("train_step_timed"))
except Exception as ex:
    print(sys.argv[2], "FAILED", ex)
PY
}
export DI_BENCH_PREDICTORS=0
: > $OUT/summary.txt
ENVV="A=1" run_bench c3_teams2 >> $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_CONV_TEAMS=1" run_bench c3_teams1 >> $OUT/summary.txt
ENVV="A=1" run_bench c5_teams2 --workload c5 --no-checks >> $OUT/summary.txt
cat $OUT/summary.txt
DEEPIMPUTE_B200_TRACE=1 DEEPIMPUTE_B200_DEEP=1 timeout 120 python scripts/trace_step.py step tf32x3 > $OUT/trace_c3_teams2.txt 2>&1
grep "trace " $OUT/trace_c3_teams2.txt | tail -4
