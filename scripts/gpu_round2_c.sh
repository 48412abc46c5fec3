#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spectral.py tests/test_gpu_scale.py::test_c2_spectral_init_residuals_at_full_size tests/test_gpu_e2e.py tests/test_gpu_api_edges.py -q --no-header -p no:cacheprovider > gpurun_out/c_pytest.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/c_pytest.log
timeout 600 python scripts/time_spectral.py > gpurun_out/c_spectral.log 2>&1
echo "spectral rc=$?"; cat gpurun_out/c_spectral.log | tail -30
