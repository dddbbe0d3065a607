#!/bin/bash
# The per-change GPU check of round 2 (run under gpurun): full GPU suite, smoke(), bench.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --maxfail=30 --tb=short -rf > gpurun_out/r02_pytest.log 2>&1; tail -6 gpurun_out/r02_pytest.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_suite.json 2> gpurun_out/r02_bench_suite.err; tail -c 300 gpurun_out/r02_bench_suite.json; tail -3 gpurun_out/r02_bench_suite.err
