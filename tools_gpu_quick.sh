#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_fusion.py -q -m gpu -x -k "dirichlet" 2>&1 | tail -5
python tools/fusion_bench.py 2>&1 | grep -E "dirichlet_fuse"
