#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_mc_launches.csv python tools/mc_probe.py 8 1 > gpurun_out/r02_mc_probe.log 2>&1
tail -2 gpurun_out/r02_mc_probe.log
