#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --maxfail=30 --tb=short -rf > gpurun_out/r02j_pytest.log 2>&1; tail -6 gpurun_out/r02j_pytest.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python tools/fit_bench.py --steps 10 2>&1 | tail -1 | tee gpurun_out/r02j_fit_bench.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; tail -c 300 gpurun_out/r02j_bench.json; tail -3 gpurun_out/r02j_bench.err
