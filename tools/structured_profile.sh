#!/bin/bash
set -u
mkdir -p gpurun_out
ncu --clock-control none --metrics gpu__time_duration.sum -c 200 --csv --log-file gpurun_out/launches_structured.csv python tools/structured_profile.py 2 > gpurun_out/structured_profile.log 2>&1
tail -3 gpurun_out/structured_profile.log
python - <<'P'
import csv,collections
rows=list(csv.reader(l for l in open('gpurun_out/launches_structured.csv') if l.startswith('"')))
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
seq=[(r[ki].split('(')[0].split('::')[-1], float(r[vi].replace(',',''))/1e6) for r in rows[1:]]
idx=[i for i,(n,t) in enumerate(seq) if n.startswith('k_sum_tables')]
a=idx[-1]
agg=collections.OrderedDict()
for n,t in seq[a:]:
    agg.setdefault(n,[0,0.0]); agg[n][0]+=1; agg[n][1]+=t
tot=sum(t for n,(c,t) in agg.items())
for n,(c,t) in sorted(agg.items(),key=lambda kv:-kv[1][1]): print(n,c,'%.3f ms %.1f%%'%(t,100*t/tot))
print('total',tot)
P
