#!/bin/bash
# Round 2, first GPU pass: parity suite with the LT kernels, pipeline traces, A/B of the new knobs on c3.
set -u
OUT=gpurun_out/r02a; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 ) > $OUT/pytest_gpu.txt
for S in 40 5; do
  ( TRACE_S=$S DEEPIMPUTE_B200_TRACE=1 timeout 120 python scripts/trace_step.py step tf32x3 2>&1 | tail -80 ) > $OUT/trace_lt_S$S.txt
done
run_bench() {  # name, env...
  name=$1; shift
  ( env "$@" timeout 400 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/bench_$name.json 2> $OUT/bench_$name.err ) 
  python - "$OUT/bench_$name.json" "$name" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d["roofline"]["kernels"]
    print(sys.argv[2], "ms_per_step %.1f e2e %.1f"%(d["ms_per_step"], d["e2e"]["ms_per_step"]), {n:(v["ms"]) for n,v in k.items()}, d["roofline"].get("predict",{}).get("ms"), d.get("engine"))
except Exception as ex:
    print(sys.argv[2], "FAILED", ex)
PY
}
export DI_BENCH_PREDICTORS=0
run_bench lt_l2 A=1 > $OUT/summary.txt
run_bench lt_nol2 DEEPIMPUTE_B200_L2_PERSIST=0 >> $OUT/summary.txt
run_bench old_l2 DEEPIMPUTE_B200_LT=0 >> $OUT/summary.txt
run_bench old_nol2 DEEPIMPUTE_B200_LT=0 DEEPIMPUTE_B200_L2_PERSIST=0 >> $OUT/summary.txt
run_bench lt_l2_tile128 DEEPIMPUTE_B200_INFER_TILE=128 >> $OUT/summary.txt
cat $OUT/summary.txt
tail -5 $OUT/pytest_gpu.txt
