#!/bin/bash
set -u
OUT=gpurun_out/r02q; mkdir -p $OUT
export DI_BENCH_PREDICTORS=0
( timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_c3.json 2> $OUT/bench_c3.err )
python - $OUT/bench_c3.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d["roofline"]; print("ms_per_step %.1f e2e %.1f"%(d["ms_per_step"], d["e2e"]["ms_per_step"]), {k:r[k] for k in ("kernel","achieved","frac","traffic","train_step_timed")}, {n:v["ms"] for n,v in r["kernels"].items()}, d.get("parity_check",{}).get("max_rel"))
PY
BENCH="python bench.py --steps 1 --warmup 0 --epochs 1 --no-cpu-baseline --no-checks"
DEEPIMPUTE_B200_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_adam' -s 40 -c 1 -o $OUT/full_c3_adam_pers_fullwidth $BENCH > $OUT/ncu1.log 2>&1
( DEEPIMPUTE_B200_LT=0 timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q -x 2>&1 | tail -3 )
