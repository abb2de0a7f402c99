#!/bin/bash
# Round 2: accumulator plan sized by chain length; API-level end to end.
set -u
OUT=gpurun_out/r02g; mkdir -p $OUT
( timeout 600 python scripts/family_accuracy.py 2>&1 | grep -v "^Epoch" | tail -30 ) > $OUT/family_accuracy.txt
cut -c1-140 $OUT/family_accuracy.txt
( timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > $OUT/pytest_gpu.txt
tail -5 $OUT/pytest_gpu.txt
run_bench() {
  name=$1; shift
  ( env $ENVV timeout 1200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err )
  python - "$OUT/bench_$name.json" "$name" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d["roofline"]["kernels"]
    pc=d.get("parity_check") or {}
    print(sys.argv[2], "ms_per_step %.1f e2e %.1f"%(d["ms_per_step"], d["e2e"]["ms_per_step"]), {n:(v["ms"]) for n,v in k.items()}, d["roofline"].get("predict",{}).get("ms"), d.get("engine"), pc.get("max_rel"), pc.get("max_rel_weights"), d["roofline"].get("train_step_timed"))
    if "api" in d: print(json.dumps(d["api"], indent=1))
except Exception as ex:
    print(sys.argv[2], "FAILED", ex)
PY
}
: > $OUT/summary.txt
ENVV="DI_BENCH_PREDICTORS=0" run_bench c3_shard8 --emulate-shard 0/8 >> $OUT/summary.txt
ENVV="DI_BENCH_PREDICTORS=0" run_bench c2 --workload c2 >> $OUT/summary.txt
ENVV="A=1" run_bench c3_api --api >> $OUT/summary.txt
cat $OUT/summary.txt
tail -3 $OUT/bench_c3_api.err
