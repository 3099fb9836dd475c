#!/bin/bash
# LPVS_PHASE_STRUCTURED_REF at cfg2: plain run, launch list, one --set full capture of k_gram_corr (gpurun_out/<tag>_*)
set -u
TAG=${1:-corr}
mkdir -p gpurun_out
export LPVS_PROFILE_MODE=5
timeout 200 python tools/structured_profile.py 3 > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
tail -2 gpurun_out/${TAG}_plain.log
timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/structured_profile.py 2 > gpurun_out/${TAG}_ncu1.log 2>&1
python - "$TAG" <<'P'
import csv,collections,sys
tag=sys.argv[1]
rows=list(csv.reader(l for l in open(f'gpurun_out/{tag}_launches.csv') if l.startswith('"')))
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
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_gram_corr -c 1 -o gpurun_out/${TAG}_gram_corr -f python tools/structured_profile.py 1 > gpurun_out/${TAG}_ncu2.log 2>&1
ls -la gpurun_out/${TAG}_gram_corr.ncu-rep
