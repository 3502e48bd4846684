#!/bin/bash
# per-launch durations of two steps (ncu, cheap metrics pass): usage gpu_launchlist.sh <tag> [ENV=..]
TAG=$1; V=${2:-X=0}
OUT=gpurun_out/$TAG; mkdir -p $OUT
env $V timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/launch_bench.log 2>&1
python - <<PY
import csv,collections
rows=list(csv.reader(open('$OUT/launches.csv', errors='ignore')))
hdr=None; agg=collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr is None or len(r)!=len(hdr): continue
    d=dict(zip(hdr,r))
    if d.get('Metric Name')!='gpu__time_duration.sum': continue
    k=d['Kernel Name'][:48]; v=float(d['Metric Value'].replace(',',''))
    u=d['Metric Unit']; v = v/1000 if u in ('ns','nsecond') else v
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,t) in agg.items(): print('%-48s %4d  total %10.1f us  avg %9.1f us'%(k,n,t,t/n))
PY
