#!/bin/bash
# ncu evidence for round 1: launch list of the bench command + full-set capture of the conv kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 96 --csv --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r01_bench_under_ncu.log 2>&1
# one stream forward (batch 16): skip the first forward's 14 conv launches (+1 head), capture the second forward's convs
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 16 -c 16 -o gpurun_out/r01_conv_igemm python tools/perf_probe.py 16 1 3 > gpurun_out/r01_ncu_full.log 2>&1
ls -la gpurun_out/
