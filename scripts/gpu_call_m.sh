#!/bin/bash
# 2 GPUs with the dependent-launch chain: scaling line + sharded-vs-single check.
tag=${1:-s3m}
out=gpurun_out/$tag; mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    scripts/multigpu_check.py > $out/multigpu_check.txt 2>&1; echo "check exit $?" >> $out/multigpu_check.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 2 --warmup 3 > $out/bench_c3_n2.json 2> $out/bench_c3_n2.err; echo "bench exit $?" >> $out/bench_c3_n2.err
tail -n 4 $out/multigpu_check.txt; python -c "
import json; d=json.load(open('$out/bench_c3_n2.json')); print('N=2 value %.4g ms/step %.1f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))"; tail -n 2 $out/bench_c3_n2.err
