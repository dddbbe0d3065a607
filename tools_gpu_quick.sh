#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r01_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r01_smoke.log
timeout 900 python bench.py --steps 50 --warmup 5 2>&1 | tail -1 > gpurun_out/r01_bench.json; cut -c1-300 gpurun_out/r01_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/r01_bench_reference.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r01_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'raw',d['e2e']['raw_dtype_inputs']['value'],'roof',d['roofline']['achieved'],d['roofline']['frac'], d['clocks'])
PY
