#!/bin/bash
# A/B of engine knobs on one workload: every variant is one short bench run (5 epochs + predict per step).
# usage (repo root on the GPU box):  bash scripts/gpu_ab.sh <tag> name1:ENV=VAL,ENV=VAL name2:ENV=VAL ...
# knobs: DEEPIMPUTE_B200_GROUPS (sub-network groups of the epoch graph), _DEEP (0/1/2 ring configs), _GRAPH=0,
#        _ADAM=ring|big, _ADAM_GROUPS (epilogue warp groups), _ADAM_STORE=tma|direct, _PDL=0|1|2, _PDL_PREFETCH=0|1,
#        _PDL_LEAD=k, _MATH=fp32|tf32|tf32x3
tag=$1; shift
out=gpurun_out/$tag; mkdir -p $out
: > $out/ab.txt
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}; [ "$envs" = "$spec" ] && envs=""
  env DI_BENCH_PREDICTORS=0 ${envs//,/ } timeout 600 python bench.py --steps 2 --warmup 1 --epochs 5 --no-cpu-baseline \
      > $out/ab_$name.json 2> $out/ab_$name.err
  python - $out/ab_$name.json $name >> $out/ab.txt <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); k = d["roofline"]["kernels"]
    print("%-16s ms/step(5 epochs+predict) %7.2f | launch-by-launch us: fwd1 %.1f fwd2 %.1f bwd %.1f adam %.1f (%.0f GB/s)" % (
        sys.argv[2], d["ms_per_step"], k["fwd1"]["ms"] * 1e3, k["fwd2"]["ms"] * 1e3, k["bwd"]["ms"] * 1e3,
        k["adam"]["ms"] * 1e3, k["adam"]["GB/s"]))
except Exception as ex:
    print(sys.argv[2], "FAILED", ex)
PY
done
cat $out/ab.txt
