#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
timeout 300 python tools/adapnet_bench.py 16 10 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_latest.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 207 -c 69 --csv --log-file gpurun_out/adapnet_launches.csv python tools/adapnet_bench.py 16 2 > gpurun_out/adapnet_ncu.log 2>&1
