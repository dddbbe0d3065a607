#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 47 -c 23 --csv --log-file gpurun_out/launches_probe.csv python tools/perf_probe.py 16 1 3 > gpurun_out/probe_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_probe.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
gi=hdr.index('Grid Size') if 'Grid Size' in hdr else None
tot=0
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    if r[ui]=='ns': v/=1e3
    tot+=v
    print('%-70s %10.1f us  grid=%s'%(r[ki][:70], v, r[gi] if gi is not None else ''))
print('total us', tot)
PY
