#!/bin/bash
# multi-GPU evidence: tools/gpu_r02_multi.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_${N}gpu.txt 2>&1
if [ "$N" = "2" ]; then
  timeout 1200 python -m pytest tests/test_gpu_multi.py -q -m gpu -rA --tb=short > gpurun_out/r02_test_gpu_multi_2gpu.log 2>&1; tail -5 gpurun_out/r02_test_gpu_multi_2gpu.log
fi
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; tail -c 800 gpurun_out/r02_bench_${N}gpu.json; tail -3 gpurun_out/r02_bench_${N}gpu.err
