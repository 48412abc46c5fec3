#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn_tc.py -q --no-header -p no:cacheprovider -x -k "pruned or oracle or scale" > gpurun_out/g_pytest.log 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/g_pytest.log | cut -c1-400
N=1000000 timeout 600 python scripts/time_knn_c4.py > gpurun_out/g_knn_1m.log 2>&1; echo "knn 1M rc=$?"; tail -4 gpurun_out/g_knn_1m.log | cut -c1-900
