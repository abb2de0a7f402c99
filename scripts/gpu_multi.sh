#!/bin/bash
# usage: gpu_multi.sh <tag> <ngpus> <workload> [steps] [warmup]   -- bench.py under torchrun on one box
tag=$1; n=$2; wl=$3; steps=${4:-2}; warm=${5:-1}
out=gpurun_out/$tag; mkdir -p $out
export DI_BENCH_PREDICTORS=0
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $n --workload $wl --steps $steps --warmup $warm > $out/bench_${wl}_n$n.json 2> $out/bench_${wl}_n$n.err
python - $out/bench_${wl}_n$n.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("n_gpus", d["n_gpus"], "ms_per_step %.1f e2e %.1f"%(d["ms_per_step"], d["e2e"]["ms_per_step"]), d["roofline"].get("train_step_timed"), d.get("engine"), d.get("sharding_check"), "h2d", d["e2e"]["h2d_bytes_per_step"], "d2h", d["e2e"]["d2h_bytes_per_step"])
except Exception as ex:
    print("FAILED", ex)
PY
tail -3 $out/bench_${wl}_n$n.err
free -g | head -2
