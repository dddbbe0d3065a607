#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k 'regex:conv_igemm_t' -s 6 -c 3 --csv --log-file gpurun_out/r02_convt_probe.csv python tools/perf_probe.py 16 1 3 > gpurun_out/r02_convt_probe.log 2>&1
grep -v "^==" gpurun_out/r02_convt_probe.csv | cut -d, -f5,10- | tail -8
