#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_fusion.py tests/test_gpu_training.py tests/test_gpu_fcn.py tests/test_gpu_models.py -q -m gpu --tb=short -rf 2>&1 | grep -v "^tests.*PASSED" > gpurun_out/r02b_pytest.log; tail -5 gpurun_out/r02b_pytest.log
