#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_layers.py tests/test_gpu_training.py -q -m gpu --tb=short -rf -s > gpurun_out/r02g_pytest.log 2>&1; tail -6 gpurun_out/r02g_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; tail -c 400 gpurun_out/r02g_bench.json; tail -3 gpurun_out/r02g_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 260 --csv --log-file gpurun_out/r02g_fit_launches.csv python tools/fit_bench.py --steps 2 --warmup 3 > gpurun_out/r02g_fit_under_ncu.log 2>&1; tail -2 gpurun_out/r02g_fit_under_ncu.log
