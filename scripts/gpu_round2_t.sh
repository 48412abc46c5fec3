#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multi.py -q --no-header -p no:cacheprovider > gpurun_out/t_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/t_pytest_multi.log | cut -c1-300
timeout 600 $TR --master-port 29721 scripts/time_tail.py > gpurun_out/t_tail2.log 2>&1; echo "tail rc=$?"; grep "us per tail\|rror" gpurun_out/t_tail2.log | tail -10
