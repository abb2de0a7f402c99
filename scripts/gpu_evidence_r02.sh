#!/bin/bash
# Round-2 evidence (one gpurun call, one GPU): sanitizer passes, ncu launch list, ncu --set full of the final kernels at
# the grids the epoch graph replays, pipeline traces.  usage: bash scripts/gpu_evidence_r02.sh <tag>
tag=${1:-r02ev}
out=gpurun_out/$tag; mkdir -p $out
export DI_BENCH_PREDICTORS=0
# ---- sanitizer
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitizer_probe.py > $out/sanitizer_$tool.txt 2>&1
  echo "exit $?" >> $out/sanitizer_$tool.txt
  tail -4 $out/sanitizer_$tool.txt
done
# ---- launch list, c3, graph mode
bash scripts/gpu_launchlist.sh $tag c3 tf32x3 > $out/launch_summary_c3.txt 2>&1
head -12 $out/launch_summary_c3.txt
# ---- full captures, c3 at 40 sub-networks per GPU: the kernels of the epoch graph at their group grids
BENCH="python bench.py --steps 1 --warmup 0 --epochs 1 --no-cpu-baseline --no-checks"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_kernel' -s 300 -c 6 -o $out/full_c3_fwdbwd_group $BENCH > $out/full_c3_fwdbwd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_adam' -s 100 -c 2 -o $out/full_c3_adam_group $BENCH > $out/full_c3_adam_group.log 2>&1
DEEPIMPUTE_B200_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_adam' -s 40 -c 1 -o $out/full_c3_adam_fullwidth $BENCH > $out/full_c3_adam_fullwidth.log 2>&1
DEEPIMPUTE_B200_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_kernel' -s 120 -c 3 -o $out/full_c3_fwdbwd_fullwidth $BENCH > $out/full_c3_fwdbwd_fullwidth.log 2>&1
# ---- the latency-bound regime: what one GPU of an 8-GPU run does (5 sub-networks): TMA-only kernels + split K
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_lt_kernel|tc_adam' -s 400 -c 8 -o $out/full_shard8_step $BENCH --emulate-shard 0/8 > $out/full_shard8.log 2>&1
bash -c "tag=$tag; out=gpurun_out/\$tag; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tc_|gather_kernel' -s 200 -c 800 --csv --log-file \$out/launches_c3shard8.csv $BENCH --emulate-shard 0/8 > \$out/launches_c3shard8.log 2>&1"
# ---- inference
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'tc_kernel|tc_lt_kernel' -c 2 -o $out/full_infer_conv python scripts/predict_only.py > $out/full_infer_conv.log 2>&1
DEEPIMPUTE_B200_LT=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'tc_kernel|tc_lt_kernel' -c 2 -o $out/full_infer_lt python scripts/predict_only.py > $out/full_infer_lt.log 2>&1
tail -2 $out/full_infer_conv.log $out/full_infer_lt.log
# ---- traces
DEEPIMPUTE_B200_TRACE=1 DEEPIMPUTE_B200_DEEP=1 timeout 120 python scripts/trace_step.py step tf32x3 > $out/trace_c3_conv.txt 2>&1
TRACE_S=5 DEEPIMPUTE_B200_TRACE=1 timeout 120 python scripts/trace_step.py step tf32x3 > $out/trace_s5_lt.txt 2>&1
grep "trace " $out/trace_c3_conv.txt $out/trace_s5_lt.txt | tail -8
ls -la $out | head -40
