#!/bin/bash
# Developer tool (GPU box): per-kernel device times of one fwd+bwd on a shape.  usage: tools/launch_times.sh [shape] [tag]
SHAPE=${1:-headline}; TAG=${2:-lt}
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}.csv python tools/prof_one.py --shape $SHAPE --reps 3 > /dev/null 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/${TAG}.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    n=r[ki].split('(')[0].replace('void ','')
    if 'b200t5' in n or 'Memset' in n or 'memset' in n: agg.setdefault(n[:64],[]).append(float(r[vi].replace(',','')))
tot=sum(sum(v)/len(v) for v in agg.values())
for n,v in agg.items(): print("%-66s n=%d avg %8.1f us min %8.1f"%(n,len(v),sum(v)/len(v)/1e3,min(v)/1e3))
print("sum of averages %.1f us"%(tot/1e3))
PY
