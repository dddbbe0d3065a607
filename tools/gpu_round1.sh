#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -40 > gpurun_out/t_all.log; tail -25 gpurun_out/t_all.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_first.json
