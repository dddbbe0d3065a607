#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --maxfail=30 --tb=short -rf 2>&1 | grep -v "PASSED" | tail -150 > gpurun_out/r02d_pytest.log; tail -12 gpurun_out/r02d_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; tail -c 600 gpurun_out/r02d_bench.json; tail -5 gpurun_out/r02d_bench.err
timeout 900 bash tools/sanitize.sh gpurun_out/sanitize_r02 > gpurun_out/r02d_sanitize.log 2>&1; cat gpurun_out/sanitize_r02/summary.txt
