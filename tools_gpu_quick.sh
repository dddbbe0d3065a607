#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_fusion.py -q -m gpu -x 2>&1 | tail -15
python tools/fusion_bench.py 2>&1 | grep -E "confusion|suffstats|dirichlet"
