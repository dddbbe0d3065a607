#!/bin/bash
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 2>&1 | tail -2
