#!/bin/bash
# bench alone on a fresh box (no pytest before it), twice: run-to-run / box-to-box spread of the tensor-bound kNN stage
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,power.limit,power.max_limit,clocks.max.sm,temperature.gpu --format=csv > gpurun_out/v_smi.txt
for i in 1 2; do
  timeout 900 python bench.py --steps 5 --warmup 3 --no-c3 --quality off > gpurun_out/v_bench$i.json 2> gpurun_out/v_bench$i.err; echo "bench$i rc=$?"
done
python - <<'PY'
import json
for i in (1, 2):
    d = json.loads(open(f"gpurun_out/v_bench{i}.json").read().strip().split("\n")[-1])
    print(i, "value", round(d["value"], 4), "e2e", round(d["e2e"]["value"], 4), d["stages"]["ms"], d["stages"]["knn_tflops"], d["clocks"])
PY
cat gpurun_out/v_smi.txt
