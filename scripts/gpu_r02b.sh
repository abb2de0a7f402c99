#!/bin/bash
# Round 2, second GPU pass: policy (LT when the state fits L2), split K, off-critical-path Philox.
set -u
OUT=gpurun_out/r02b; mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 ) > $OUT/pytest_gpu.txt
for K in 1 4; do
  ( DEEPIMPUTE_B200_LT=1 DEEPIMPUTE_B200_SPLITK=$K TRACE_S=5 DEEPIMPUTE_B200_TRACE=1 timeout 120 python scripts/trace_step.py step tf32x3 2>&1 | tail -80 ) > $OUT/trace_lt_S5_k$K.txt
done
run_bench() {  # name, bench args..., env via ENVV
  name=$1; shift
  ( env $ENVV timeout 400 python bench.py --steps 1 --warmup 1 --no-cpu-baseline "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err )
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
: > $OUT/summary.txt
ENVV="A=1" run_bench c3_auto >> $OUT/summary.txt
ENVV="A=1" run_bench c3_shard8_auto --emulate-shard 0/8 >> $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_SPLITK=1" run_bench c3_shard8_k1 --emulate-shard 0/8 >> $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_SPLITK=2" run_bench c3_shard8_k2 --emulate-shard 0/8 >> $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_LT=0" run_bench c3_shard8_old --emulate-shard 0/8 >> $OUT/summary.txt
ENVV="A=1" run_bench c3_shard4_auto --emulate-shard 0/4 >> $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_LT=0" run_bench c3_shard4_old --emulate-shard 0/4 >> $OUT/summary.txt
ENVV="A=1" run_bench c3_shard2_auto --emulate-shard 0/2 >> $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_LT=0" run_bench c3_shard2_old --emulate-shard 0/2 >> $OUT/summary.txt
ENVV="A=1" run_bench c2_auto --workload c2 >> $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_LT=0" run_bench c2_old --workload c2 >> $OUT/summary.txt
cat $OUT/summary.txt
tail -5 $OUT/pytest_gpu.txt
