#!/bin/bash
# round 2, first GPU pass: full GPU suite, smoke, bench, fusion-kernel ncu capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r02a_gpu.txt 2>&1
timeout 2400 python -m pytest tests -q -m gpu --maxfail=20 -rf 2>&1 | tail -80 > gpurun_out/r02a_pytest.log; tail -30 gpurun_out/r02a_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r02a_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; tail -c 3000 gpurun_out/r02a_bench.json; tail -5 gpurun_out/r02a_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dirichlet_fuse|softmax_argmax|suffstats|confusion|bayes_lut" -c 14 -o gpurun_out/r02a_fusion python tools/fusion_bench.py > gpurun_out/r02a_fusion_ncu.log 2>&1
ls -la gpurun_out | tail -12
