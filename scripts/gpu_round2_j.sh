#!/bin/bash
# 2-GPU pass: push tail parity + timing, bench N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
MMUMAP_PEER_TAIL=push timeout 600 $TR --master-port 29651 scripts/check_multigpu.py > gpurun_out/j_parity_push.log 2>&1; echo "push parity rc=$?"; grep -E "exchange|epoch|OK|rror" gpurun_out/j_parity_push.log | tail -8
timeout 600 $TR --master-port 29652 scripts/time_tail.py > gpurun_out/j_tail2.log 2>&1; echo "tail rc=$?"; grep "us per tail\|rror" gpurun_out/j_tail2.log | tail -8
timeout 900 $TR --master-port 29653 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline --quality off --no-c3 --no-transform > gpurun_out/j_bench2.json 2> gpurun_out/j_bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/j_bench2.json").read().strip().split("\n")[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["config"]["epoch_tail_kernel"], d["stages"]["ms"], d["stages"]["epoch_kernels_us_per_launch"])
PY
N=4000000 MMUMAP_KNN_DEBUG=1 timeout 900 $TR --master-port 29654 scripts/time_knn_c4_dist.py > gpurun_out/j_knn4m_2gpu.log 2>&1; echo "knn 4M x2 rc=$?"; grep -E "knn_pruned|ms " gpurun_out/j_knn4m_2gpu.log | tail -6 | cut -c1-1200
