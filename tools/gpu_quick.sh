#!/bin/bash
timeout 500 python -m pytest tests/test_gpu_models.py tests/test_gpu_fullsize.py -q -m gpu -x 2>&1 | tail -3
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
