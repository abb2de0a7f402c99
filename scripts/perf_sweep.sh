#!/bin/bash
# Scratch: sweep the epoch-graph knobs on one workload.  usage: perf_sweep.sh <tag> <workload> <math>
tag=$1; wl=$2; m=$3; out=gpurun_out/$tag; mkdir -p $out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 900 python bench.py --workload $wl --math $m --steps 2 --warmup 1 --epochs 5 --no-cpu-baseline > $out/$name.json 2> $out/$name.err
  python - "$out/$name.json" "$name" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print("%-28s ms/epoch %8.2f  value %.3e"%(sys.argv[2], d["ms_per_step"]/5, d["value"]))
except Exception as ex:
    print(sys.argv[2], "FAILED", ex)
PY
}
run nograph DEEPIMPUTE_B200_GRAPH=0
run g1_shallow DEEPIMPUTE_B200_GROUPS=1 DEEPIMPUTE_B200_DEEP=0
run g2_shallow DEEPIMPUTE_B200_GROUPS=2 DEEPIMPUTE_B200_DEEP=0
run g4_shallow DEEPIMPUTE_B200_GROUPS=4 DEEPIMPUTE_B200_DEEP=0
run g4_deep DEEPIMPUTE_B200_GROUPS=4 DEEPIMPUTE_B200_DEEP=1
run g8_shallow DEEPIMPUTE_B200_GROUPS=8 DEEPIMPUTE_B200_DEEP=0
run g8_deep DEEPIMPUTE_B200_GROUPS=8 DEEPIMPUTE_B200_DEEP=1
tail -n 3 $out/*.err | head -40
