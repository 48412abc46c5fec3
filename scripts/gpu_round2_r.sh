#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_layout.py -q --no-header -p no:cacheprovider -x > gpurun_out/r_pytest_layout.log 2>&1; echo "layout tests rc=$?"; tail -3 gpurun_out/r_pytest_layout.log | cut -c1-200
timeout 400 python scripts/time_layout_c4.py 10000000 80 > gpurun_out/r_c4_layout.log 2>&1; tail -2 gpurun_out/r_c4_layout.log
