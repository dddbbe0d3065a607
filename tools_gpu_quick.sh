#!/bin/bash
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:decode_upsample8' -c 3 --csv --log-file gpurun_out/tail_launches.csv python tools/perf_probe.py 16 1 3 > gpurun_out/tail_ncu.log 2>&1
grep -E "decode" gpurun_out/tail_launches.csv | awk -F'","' '{print $NF}'
timeout 900 python -m pytest tests/test_gpu_fcn.py tests/test_gpu_models.py tests/test_gpu_fullsize.py -q -m gpu -x 2>&1 | tail -3
