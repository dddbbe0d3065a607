#!/bin/bash
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_igemm -c 45 --csv --log-file gpurun_out/c1.csv python tools/perf_probe.py 16 1 3 > /dev/null 2>&1
grep -E "64, 1, 0, 3|igemm_t_kernel<1>" gpurun_out/c1.csv | awk -F'","' '{print $5, $(NF)}' | tr -d '"' | head -8
