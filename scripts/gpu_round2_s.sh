#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29711 scripts/time_tail.py > gpurun_out/s_tail2.log 2>&1; echo "tail rc=$?"; grep "us per tail\|rror" gpurun_out/s_tail2.log | tail -10
