#!/bin/bash
# usage: gpu_check_r02.sh <tag>   -- the lean round-end check: GPU tests, smoke, the default bench line (no profiler passes)
tag=${1:-r02check}; out=gpurun_out/$tag; mkdir -p $out
timeout 600 python -m pytest tests -x -q -m gpu > $out/pytest_gpu.txt 2>&1; tail -2 $out/pytest_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" > $out/smoke.txt 2>&1; tail -1 $out/smoke.txt
timeout 600 python bench.py > $out/bench_c3.json 2> $out/bench_c3.err
python - $out/bench_c3.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("c3 ms_per_step %.1f e2e %.1f frac %.3f timed %s parity %s fallbacks %s"%(d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["train_step_timed"], d.get("parity_check",{}).get("max_rel"), d.get("graph_fallbacks")))
    print(json.dumps(d["config"])); print(json.dumps(d["arm"]))
except Exception as ex:
    print("FAILED", ex)
PY
tail -2 $out/bench_c3.err
