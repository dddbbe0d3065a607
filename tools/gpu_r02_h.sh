#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_training.py tests/test_gpu_layers.py -q -m gpu --tb=short -rf > gpurun_out/r02h_pytest.log 2>&1; tail -6 gpurun_out/r02h_pytest.log | cut -c1-300
timeout 600 python tools/fit_bench.py --steps 10 2>&1 | tail -1 | tee gpurun_out/r02h_fit_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 420 -c 140 --csv --log-file gpurun_out/r02h_fit_launches.csv python tools/fit_bench.py --steps 2 --warmup 3 > gpurun_out/r02h_fit_under_ncu.log 2>&1
