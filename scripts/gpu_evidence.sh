#!/bin/bash
# Full evidence run (one gpurun call): GPU tests, smoke, bench (both arms, c3 and c2), ncu launch list + full captures of the
# ADAM, element-wise and inference kernels, pipeline trace.  usage: bash scripts/gpu_evidence.sh <tag>; then
# python scripts/summarise_profile.py gpurun_out/<tag> <round-tag>
tag=${1:-evidence}
out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q --durations=5 > $out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1; echo "smoke exit $?" >> $out/smoke.txt
timeout 900 python bench.py > $out/bench_c3.json 2> $out/bench_c3.err
timeout 600 python bench.py --workload c2 > $out/bench_c2.json 2> $out/bench_c2.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err
bash scripts/gpu_launchlist.sh $tag c3 tf32x3 > $out/launch_summary.txt 2>&1
DEEPIMPUTE_B200_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_adam -s 40 -c 1 -o $out/full_c3_adam \
    python bench.py --steps 1 --warmup 0 --epochs 1 --no-cpu-baseline > $out/full_c3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'impute_kernel|counts_to_norm' -c 3 -o $out/full_post_impute \
    python scripts/trace_step.py impute > $out/full_impute.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_kernel -c 2 -o $out/full_infer_tckernel \
    python scripts/predict_only.py > $out/full_infer.log 2>&1
DEEPIMPUTE_B200_TRACE=1 DEEPIMPUTE_B200_DEEP=1 timeout 120 python scripts/trace_step.py step tf32x3 > $out/trace_x3.txt 2>&1
tail -n 8 $out/pytest_gpu.txt; cat $out/smoke.txt; cat $out/bench_c3.json $out/bench_c2.json; tail -n 2 $out/*.err; cat $out/launch_summary.txt; grep "trace " $out/trace_x3.txt | tail -4
