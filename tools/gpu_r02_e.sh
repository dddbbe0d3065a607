#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_training.py -q -m gpu --tb=short -rf -s > gpurun_out/r02e_training.log 2>&1; tail -5 gpurun_out/r02e_training.log
timeout 1200 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_fcn.py tests/test_gpu_models.py -q -m gpu --tb=short -rf -s > gpurun_out/r02e_full.log 2>&1; tail -8 gpurun_out/r02e_full.log
