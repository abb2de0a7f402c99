#!/bin/bash
# Full evidence run: GPU tests, smoke, bench lines (both arms), ncu launch list + full capture of the dominant kernel.
tag=${1:-final}
out=gpurun_out/$tag; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q > $out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1; echo "smoke exit $?" >> $out/smoke.txt
timeout 1500 python bench.py > $out/bench_c3.json 2> $out/bench_c3.err
timeout 900 python bench.py --workload c2 > $out/bench_c2.json 2> $out/bench_c2.err
timeout 900 python bench.py --math tf32 > $out/bench_c3_tf32.json 2> $out/bench_c3_tf32.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err
tail -n 4 $out/pytest_gpu.txt; cat $out/smoke.txt; cat $out/bench_c3.json $out/bench_c2.json $out/bench_c3_tf32.json $out/bench_ref.json; tail -n 3 $out/*.err
