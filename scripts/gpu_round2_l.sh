#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multi.py -q --no-header -p no:cacheprovider > gpurun_out/l_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -4 gpurun_out/l_pytest_multi.log | cut -c1-300
timeout 600 $TR --master-port 29662 scripts/time_tail.py > gpurun_out/l_tail2.log 2>&1; echo "tail rc=$?"; grep "us per tail\|rror" gpurun_out/l_tail2.log | tail -8
timeout 900 $TR --master-port 29663 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline --quality device > gpurun_out/l_bench2.json 2> gpurun_out/l_bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/l_bench2.json").read().strip().split("\n")[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["config"]["epoch_tail_kernel"], d["stages"]["ms"], d["stages"]["epoch_kernels_us_per_launch"], "c3", d["stages"]["c3"], "q", d["quality"])
PY
tail -2 gpurun_out/l_bench2.err | cut -c1-300
