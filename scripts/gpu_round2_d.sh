#!/bin/bash
# 2-GPU pass: parity of both kNN modes and the optimiser exchange forms, then the bench at N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/d_topo.txt 2>&1
timeout 600 $TR --master-port 29611 scripts/check_multigpu.py > gpurun_out/d_parity_fused.log 2>&1; echo "fused rc=$?"; grep -E "exchange|epoch|OK|Error|error" gpurun_out/d_parity_fused.log | tail -8
MMUMAP_PEER_MULTIMEM=0 timeout 600 $TR --master-port 29612 scripts/check_multigpu.py > gpurun_out/d_parity_fused_nomc.log 2>&1; echo "fused-nomc rc=$?"; grep -E "exchange|epoch|OK|Error|error" gpurun_out/d_parity_fused_nomc.log | tail -8
MMUMAP_PEER_TAIL=legacy timeout 600 $TR --master-port 29613 scripts/check_multigpu.py > gpurun_out/d_parity_legacy.log 2>&1; echo "legacy rc=$?"; grep -E "exchange|epoch|OK|Error|error" gpurun_out/d_parity_legacy.log | tail -8
timeout 900 $TR --master-port 29614 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/d_bench2.json 2> gpurun_out/d_bench2.err; echo "bench2 rc=$?"
tail -c 2500 gpurun_out/d_bench2.json; tail -5 gpurun_out/d_bench2.err
