#!/bin/bash
# first GPU pass: fusion kernels, single layers, whole expert
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_fusion.py -q -m gpu -x 2>&1 | tail -40 > gpurun_out/t_fusion.log
timeout 900 python -m pytest tests/test_gpu_layers.py -q -m gpu 2>&1 | tail -150 > gpurun_out/t_layers.log
timeout 900 python -m pytest tests/test_gpu_fcn.py -q -m gpu -s 2>&1 | tail -150 > gpurun_out/t_fcn.log
tail -5 gpurun_out/t_fusion.log; tail -30 gpurun_out/t_layers.log; tail -30 gpurun_out/t_fcn.log
