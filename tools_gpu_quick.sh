#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_fullsize.py tests/test_gpu_adapnet.py -q -m gpu -x 2>&1 | tail -5
timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_latest.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_latest.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'raw',d['e2e']['raw_dtype_inputs']['value'],'roof',d['roofline']['achieved'],d['roofline']['frac'], d['clocks'])
PY
