#!/bin/bash
# round 2, pass C: full GPU suite, smoke, bench, launch list of one bench step
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --maxfail=30 --tb=short -rf 2>&1 | grep -v "PASSED" | tail -150 > gpurun_out/r02c_pytest.log; tail -25 gpurun_out/r02c_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r02c_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; tail -c 1500 gpurun_out/r02c_bench.json; tail -5 gpurun_out/r02c_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_bench_launches.csv python bench.py --steps 2 --warmup 3 --soak 0 --no-extras > gpurun_out/r02c_bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -8
