#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_layers.py tests/test_gpu_fcn.py tests/test_gpu_models.py -q -m gpu -x 2>&1 | tail -15
python tools/perf_probe.py 16 5 3 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -s 41 -c 20 --csv --log-file gpurun_out/launches_probe.csv python tools/perf_probe.py 16 1 3 > gpurun_out/probe_ncu.log 2>&1
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
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'conv TF',round(d['roofline']['achieved'],1),'ms',round(d['ms_per_step'],3),'e2e ms',round(d['e2e']['ms_per_step'],3))"
