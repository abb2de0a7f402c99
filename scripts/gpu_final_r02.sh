#!/bin/bash
# Final state of round 2 on one GPU: what the driver runs (tests, smoke, bench both arms) + the ncu captures of the
# persistent ADAM kernel.
tag=${1:-r02final}
out=gpurun_out/$tag; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > $out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.txt
tail -5 $out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1; echo "smoke exit $?" >> $out/smoke.txt; cat $out/smoke.txt | tail -5
timeout 900 python bench.py > $out/bench_c3.json 2> $out/bench_c3.err; tail -c 600 $out/bench_c3.json; echo
timeout 900 python bench.py --impl reference > $out/bench_c3_ref.json 2> $out/bench_c3_ref.err; tail -c 400 $out/bench_c3_ref.json; echo
export DI_BENCH_PREDICTORS=0
BENCH="python bench.py --steps 1 --warmup 0 --epochs 1 --no-cpu-baseline --no-checks"
DEEPIMPUTE_B200_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_adam' -s 40 -c 1 -o $out/full_c3_adam_pers_fullwidth $BENCH > $out/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_adam' -s 100 -c 2 -o $out/full_c3_adam_pers_group $BENCH > $out/ncu2.log 2>&1
bash scripts/gpu_launchlist.sh $tag c3 tf32x3 > $out/launch_summary_c3.txt 2>&1; head -10 $out/launch_summary_c3.txt
timeout 900 python bench.py --workload c2 --no-cpu-baseline > $out/bench_c2.json 2> $out/bench_c2.err; tail -c 300 $out/bench_c2.json; echo
