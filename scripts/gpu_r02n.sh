#!/bin/bash
# Persistent inference kernel: parity suite, speed, tensor-pipe.
set -u
OUT=gpurun_out/r02n; mkdir -p $OUT
( timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 ) > $OUT/pytest_gpu.txt
tail -6 $OUT/pytest_gpu.txt
( timeout 300 python scripts/predict_only.py 2>&1 | tail -3 ) > $OUT/predict_only.txt; cat $OUT/predict_only.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'tc_lt_infer_kernel' -c 2 -o $OUT/full_infer_persistent python scripts/predict_only.py > $OUT/full_infer_persistent.log 2>&1
ncu -i $OUT/full_infer_persistent.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]:
    g=lambda k: r[h.index(k)] if k in h else '-'
    print(g('Kernel Name')[:40], g('Grid Size'), 't', g('gpu__time_duration.sum'), 'tensor%', g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'), 'dram%', g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'), 'rd', g('dram__bytes_read.sum'), 'wr', g('dram__bytes_write.sum'))
"
export DI_BENCH_PREDICTORS=0
( timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/bench_c3.json 2> $OUT/bench_c3.err )
python - $OUT/bench_c3.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k=d["roofline"]["kernels"]; pc=d.get("parity_check") or {}
print("c3 ms_per_step %.1f e2e %.1f"%(d["ms_per_step"], d["e2e"]["ms_per_step"]), {n:(v["ms"]) for n,v in k.items()}, d["roofline"].get("predict"), pc.get("max_rel"), pc.get("max_rel_weights"))
PY
