#!/bin/bash
# Round 2, third GPU pass: early dependent-launch release + cheaper split-K exchange; c5; bench checks; reference arm.
set -u
OUT=gpurun_out/r02c; mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > $OUT/pytest_gpu.txt
run_bench() {  # name, bench args..., env via ENVV
  name=$1; shift
  ( env $ENVV timeout 900 python bench.py --steps 1 --warmup 1 "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err )
  python - "$OUT/bench_$name.json" "$name" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d["roofline"]["kernels"]
    print(sys.argv[2], "ms_per_step %.1f e2e %.1f"%(d["ms_per_step"], d["e2e"]["ms_per_step"]), {n:(v["ms"]) for n,v in k.items()}, d["roofline"].get("predict",{}).get("ms"), d.get("engine"), d.get("parity_check"), d["roofline"].get("train_step_timed"), d.get("cpu_baseline",{}).get("value"))
except Exception as ex:
    print(sys.argv[2], "FAILED", ex)
PY
}
export DI_BENCH_PREDICTORS=0
: > $OUT/summary.txt
ENVV="A=1" run_bench c3_shard8 --emulate-shard 0/8 --no-cpu-baseline >> $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_PDL=1" run_bench c3_shard8_late --emulate-shard 0/8 --no-cpu-baseline --no-checks >> $OUT/summary.txt
ENVV="A=1" run_bench c3_shard4 --emulate-shard 0/4 --no-cpu-baseline --no-checks >> $OUT/summary.txt
ENVV="A=1" run_bench c3_shard2 --emulate-shard 0/2 --no-cpu-baseline --no-checks >> $OUT/summary.txt
ENVV="DEEPIMPUTE_B200_PDL=2" run_bench c3_shard2_early --emulate-shard 0/2 --no-cpu-baseline --no-checks >> $OUT/summary.txt
ENVV="A=1" run_bench c2 --workload c2 --no-cpu-baseline >> $OUT/summary.txt
ENVV="A=1" run_bench c3_full >> $OUT/summary.txt
ENVV="A=1" run_bench c5 --workload c5 --no-cpu-baseline >> $OUT/summary.txt
( timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $OUT/bench_ref_c3.json 2> $OUT/bench_ref_c3.err ); tail -c 1500 $OUT/bench_ref_c3.json >> $OUT/summary.txt
cat $OUT/summary.txt
tail -5 $OUT/pytest_gpu.txt
tail -5 $OUT/bench_c5.err
free -g | head -2
