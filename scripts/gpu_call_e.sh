#!/bin/bash
# Epilogue warp groups of the one-CTA-per-SM ADAM kernel.
tag=${1:-s3e}
out=gpurun_out/$tag; mkdir -p $out
timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_multinet_gpu.py -m gpu -q > $out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.txt
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
run g4 DEEPIMPUTE_B200_ADAM_GROUPS=4 > $out/ab.txt
run g3 DEEPIMPUTE_B200_ADAM_GROUPS=3 >> $out/ab.txt
run g2 DEEPIMPUTE_B200_ADAM_GROUPS=2 >> $out/ab.txt
DEEPIMPUTE_B200_ADAM_GROUPS=4 DEEPIMPUTE_B200_TRACE=1 DEEPIMPUTE_B200_DEEP=1 timeout 120 python scripts/trace_step.py step tf32x3 > $out/trace_x3_g4.txt 2>&1
tail -n 4 $out/pytest_gpu.txt; cat $out/ab.txt; grep -A20 "trace adam" $out/trace_x3_g4.txt | tail -n 22
