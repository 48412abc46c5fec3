#!/bin/bash
# the driver's own sequence on one GPU with the final tree: smoke(), then bench.py with its default flags
mkdir -p gpurun_out
T0=$SECONDS; timeout 900 python bench.py > gpurun_out/y_bench_default.json 2> gpurun_out/y_bench_default.err; echo "bench rc=$? wall $((SECONDS-T0)) s"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/y_bench_default.json").read().strip().split("\n")[-1])
print("steps", d["steps"], "warmup", d["warmup"], "value", round(d["value"], 4), "e2e", round(d["e2e"]["value"], 4), d["e2e"]["seconds_each_step_rank0"])
print(d["stages"]["ms"], d["stages"]["knn_ms_each_call"])
print("roofline", d["roofline"]["frac"], d["roofline"]["l2_roof"]["frac"], "knn", d["roofline_other"]["frac"], d["clocks"])
print("quality", json.dumps(d["quality"])[:600])
print("c3", d["stages"]["c3"]["fit_s"], "transform", d["stages"]["transform_100k"]["seconds"], "cpu_baseline", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"])
PY
