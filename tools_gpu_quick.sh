#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 160 --csv --log-file gpurun_out/launches_train.csv python tools/train_probe.py 16 1 > gpurun_out/train_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_train.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.OrderedDict(); tot=0
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    if r[ui]=='ns': v/=1e3
    name=r[ki].replace('void ','').replace('<unnamed>::','').split('(')[0][:55]
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v; tot+=v
print('launches',sum(a[0] for a in agg.values()),'total us',round(tot,1))
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:25]:
    print('%-57s n=%3d  %9.1f us  %5.1f%%'%(k,n,t,100*t/tot))
PY
