#!/bin/bash
# usage: gpu_launchlist.sh <tag> <workload> <math> [ENV=VAL ...]  -- ncu launch list (device time per launch)
tag=$1; wl=$2; m=$3; shift 3
out=gpurun_out/$tag; mkdir -p $out
env "$@" timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tc_|gather_kernel|simt_' -s 200 -c 1200 \
    --csv --log-file $out/launches_$wl.csv python bench.py --workload $wl --math $m --steps 1 --warmup 0 --epochs 1 --no-cpu-baseline > $out/launches_$wl.log 2>&1
python - $out/launches_$wl.csv <<'PY'
import csv, sys, collections
lines=[l for l in open(sys.argv[1]) if l.startswith('"')]
rows=list(csv.DictReader(lines))
agg=collections.defaultdict(list)
for r in rows:
    name=r['Kernel Name']
    short=name.split('(')[0].split('::')[-1]
    agg[short+' grid='+r['Grid Size']].append(float(r['Metric Value']))
tot=sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
    print("%-60s n=%4d mean=%8.1f ns share=%.3f"%(k,len(v),sum(v)/len(v),sum(v)/tot))
PY
