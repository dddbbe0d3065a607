#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/e2e_split_probe.py 2>&1 | tail -6 | tee gpurun_out/r02i_split_probe.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02i_fit_launches.csv python tools/fit_bench.py --steps 1 --warmup 1 > gpurun_out/r02i_fit_under_ncu.log 2>&1
