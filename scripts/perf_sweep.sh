#!/bin/bash
# Scratch: sweep the epoch-graph knobs on one workload.  usage: perf_sweep.sh <tag> <workload> <math> [variants...]
tag=$1; wl=$2; m=$3; shift 3; out=gpurun_out/$tag; mkdir -p $out
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
for v in "$@"; do
  case $v in
    nograph) run nograph DEEPIMPUTE_B200_GRAPH=0 ;;
    g*_deep) g=${v#g}; g=${g%_deep}; run $v DEEPIMPUTE_B200_GROUPS=$g DEEPIMPUTE_B200_DEEP=1 ;;
    g*_medium) g=${v#g}; g=${g%_medium}; run $v DEEPIMPUTE_B200_GROUPS=$g DEEPIMPUTE_B200_DEEP=2 ;;
    g*_shallow) g=${v#g}; g=${g%_shallow}; run $v DEEPIMPUTE_B200_GROUPS=$g DEEPIMPUTE_B200_DEEP=0 ;;
  esac
done
tail -qn 2 $out/*.err | head -20
