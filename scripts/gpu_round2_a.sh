#!/bin/bash
# first GPU pass of round 2: tests, bench, launch list, C4-shaped optimiser timing
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt
timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 300 python scripts/time_layout_c4.py 10000000 0 48 32 > gpurun_out/a_c4_layout.log 2>&1
tail -4 gpurun_out/a_c4_layout.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/a_bench.json
