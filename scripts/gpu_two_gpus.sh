#!/bin/bash
# 2 GPUs: sharded MultiNet against the single-GPU one, and the bench line the driver would ask for at N=2.
tag=${1:-two_gpus}
out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv > $out/gpus.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    scripts/multigpu_check.py > $out/multigpu_check.txt 2>&1; echo "check exit $?" >> $out/multigpu_check.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 2 --warmup 3 > $out/bench_c3_n2.json 2> $out/bench_c3_n2.err; echo "bench exit $?" >> $out/bench_c3_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > $out/bench_ref_n2.json 2> $out/bench_ref_n2.err
cat $out/gpus.txt; tail -n 12 $out/multigpu_check.txt; cat $out/bench_c3_n2.json; tail -n 4 $out/bench_c3_n2.err; cat $out/bench_ref_n2.json | cut -c1-300
