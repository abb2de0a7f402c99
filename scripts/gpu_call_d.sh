#!/bin/bash
# One-CTA-per-SM ADAM kernel against the ring kernel, GPU tests, traces, UMMA micro-benchmark, ncu captures.
tag=${1:-s3d}
out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q --durations=5 > $out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.txt
run() { name=$1; shift
  env DI_BENCH_PREDICTORS=0 "$@" timeout 600 python bench.py --steps 2 --warmup 1 --epochs 5 --no-cpu-baseline > $out/ab_$name.json 2> $out/ab_$name.err
  python - $out/ab_$name.json $name <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); k=d["roofline"]["kernels"]
    print("%-14s ms/step(5 epochs+predict) %7.2f | launch-by-launch us: fwd1 %.1f fwd2 %.1f bwd %.1f adam %.1f (%.0f GB/s)" % (sys.argv[2], d["ms_per_step"],
          k["fwd1"]["ms"]*1e3, k["fwd2"]["ms"]*1e3, k["bwd"]["ms"]*1e3, k["adam"]["ms"]*1e3, k["adam"]["GB/s"]))
except Exception as ex:
    print(sys.argv[2], "FAILED", ex)
PY
}
run big   DEEPIMPUTE_B200_ADAM=big > $out/ab.txt
run ring  DEEPIMPUTE_B200_ADAM=ring >> $out/ab.txt
run big_g8   DEEPIMPUTE_B200_ADAM=big DEEPIMPUTE_B200_GROUPS=8 >> $out/ab.txt
run big_g40  DEEPIMPUTE_B200_ADAM=big DEEPIMPUTE_B200_GROUPS=40 >> $out/ab.txt
DEEPIMPUTE_B200_TRACE=1 DEEPIMPUTE_B200_DEEP=1 timeout 120 python scripts/trace_step.py step tf32x3 > $out/trace_x3_big.txt 2>&1
timeout 120 deepimpute_b200/csrc/umma_bench > $out/umma_bench.txt 2>&1
DEEPIMPUTE_B200_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_adam -s 40 -c 1 -o $out/full_c3_adam \
    python bench.py --steps 1 --warmup 0 --epochs 1 --no-cpu-baseline > $out/full_c3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'impute_kernel|counts_to_norm' -c 3 -o $out/full_post_impute \
    python scripts/trace_step.py impute > $out/full_impute.log 2>&1
tail -n 8 $out/pytest_gpu.txt; cat $out/ab.txt; cat $out/umma_bench.txt; grep -A22 "trace adam" $out/trace_x3_big.txt | tail -n 24; tail -n 2 $out/full_c3.log | cut -c1-200; ls -la $out
