#!/bin/bash
# round-1 evidence: full GPU test suite, smoke, bench, ncu launch list + full-set capture
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r01_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r01_smoke.log
timeout 900 python bench.py --steps 50 --warmup 5 2>&1 | tail -1 > gpurun_out/r01_bench.json; cat gpurun_out/r01_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/r01_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r01_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:conv_igemm|conv_c1' -s 15 -c 15 -o gpurun_out/r01_conv_igemm python tools/perf_probe.py 16 1 3 > gpurun_out/r01_ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:"dirichlet_fuse|softmax_argmax|mean_fuse|confusion|mc_moments" -c 12 -o gpurun_out/r01_fusion python tools/fusion_bench.py > gpurun_out/r01_fusion_ncu.log 2>&1
python tools/fusion_bench.py > gpurun_out/r01_fusion_roofline.txt 2>&1; cp gpurun_out/fusion_roofline.json gpurun_out/r01_fusion_roofline.json
python tools/timing.py --repetitions 30 --graph --cpu --json gpurun_out/r01_timing_sweep.json > gpurun_out/r01_timing_sweep.txt 2>&1
python tools/fit_bench.py --steps 10 > gpurun_out/r01_fit_bench.json 2>&1
timeout 300 python tools/adapnet_bench.py 16 10 2>&1 | tail -1 > gpurun_out/r01_adapnet_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 207 -c 69 --csv --log-file gpurun_out/r01_adapnet_launches.csv python tools/adapnet_bench.py 16 2 > gpurun_out/r01_adapnet_ncu.log 2>&1
ls -la gpurun_out | tail -20
