#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r01_bench_under_ncu.log 2>&1
tail -2 gpurun_out/r01_bench_under_ncu.log | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_fcn.py -q -m gpu -x -k variants 2>&1 | tail -3
