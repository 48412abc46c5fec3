#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29631 scripts/time_tail.py > gpurun_out/f_tail.log 2>&1; echo "tail rc=$?"; grep "us per tail\|Error\|error" gpurun_out/f_tail.log | tail -12
timeout 900 $TR --master-port 29632 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline --quality off --no-c3 --no-transform > gpurun_out/f_bench2.json 2> gpurun_out/f_bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/f_bench2.json").read().strip().split("\n")[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["stages"]["ms"], d["stages"]["epoch_kernels_us_per_launch"])
PY
