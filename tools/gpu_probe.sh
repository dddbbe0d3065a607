#!/bin/bash
mkdir -p gpurun_out
python tools/perf_probe.py 16 5 3 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -s 47 -c 23 --csv --log-file gpurun_out/launches_probe.csv python tools/perf_probe.py 16 1 3 > gpurun_out/probe_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_probe.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
tot=0
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    if r[ui]=='ns': v/=1e3
    tot+=v
    print('%-60s %10.1f us'%(r[ki].replace('void ','').replace('<unnamed>::','')[:60], v))
print('total us', tot)
PY
