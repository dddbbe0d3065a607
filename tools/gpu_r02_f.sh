#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --maxfail=30 --tb=short -rf > gpurun_out/r02f_pytest.log 2>&1; tail -12 gpurun_out/r02f_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
