#!/bin/bash
timeout 200 python tools/halo_bench.py 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_layers.py tests/test_gpu_fcn.py tests/test_gpu_adapnet.py tests/test_gpu_training.py -q -m gpu -x 2>&1 | tail -5
timeout 300 python tools/adapnet_bench.py 16 10 2>&1 | tail -1
