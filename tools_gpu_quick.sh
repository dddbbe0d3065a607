#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_fcn.py -q -m gpu -x 2>&1 | tail -15
