#!/bin/bash
timeout 120 python tools/conv_bench.py 0,2048 conv3_1,conv3_2,conv4_2,conv5_1 2>&1 | tail -10
timeout 300 python -m pytest tests/test_gpu_layers.py -q -m gpu -x 2>&1 | tail -5
