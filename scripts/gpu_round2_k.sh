#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn_tc.py tests/test_gpu_scale.py::test_c4_shaped_slice_knn_bit_exact tests/test_gpu_spectral.py -q --no-header -p no:cacheprovider -x > gpurun_out/k_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/k_pytest.log | cut -c1-400
N=1000000 MMUMAP_KNN_DEBUG=1 timeout 600 python scripts/time_knn_c4.py > gpurun_out/k_knn_1m.log 2>&1; echo "knn 1M rc=$?"; grep -E "knn_pruned|ms " gpurun_out/k_knn_1m.log | tail -3 | cut -c1-1300
N=10000000 MMUMAP_KNN_DEBUG=1 timeout 900 python scripts/time_knn_c4.py > gpurun_out/k_knn_10m.log 2>&1; echo "knn 10M rc=$?"; grep -E "knn_pruned|ms " gpurun_out/k_knn_10m.log | tail -3 | cut -c1-1300
timeout 600 python scripts/time_invert.py > gpurun_out/k_invert.log 2>&1; echo "invert rc=$?"; tail -2 gpurun_out/k_invert.log
