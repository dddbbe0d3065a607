#!/bin/bash
timeout 300 python tools/c1_debug.py 3 2>&1 | tail -1
timeout 300 python tools/c1_debug.py 1 2>&1 | tail -1
timeout 300 python tools/conv_bench.py 0,1,8,13 conv1_1 2>&1 | tail -12
timeout 900 python -m pytest tests/test_gpu_fcn.py tests/test_gpu_layers.py tests/test_gpu_adapnet.py -q -m gpu -x 2>&1 | tail -3
