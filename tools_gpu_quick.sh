#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 50 --warmup 5 2>&1 | tail -1 > gpurun_out/r01_bench.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r01_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'raw',d['e2e']['raw_dtype_inputs']['value'],'roof',d['roofline']['achieved'],d['roofline']['frac'], d['clocks'])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r01_bench_under_ncu.log 2>&1
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
