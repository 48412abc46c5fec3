#!/bin/bash
# bench alone, three times on one box: per-call kNN stage times, warm-up with the timed loop's model lifetime
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 900 python bench.py --steps 5 --warmup 3 --no-c3 --quality off > gpurun_out/z_bench$i.json 2> gpurun_out/z_bench$i.err; echo "bench$i rc=$?"
done
python - <<'PY'
import json
for i in (1, 2, 3):
    d = json.loads(open(f"gpurun_out/z_bench{i}.json").read().strip().split("\n")[-1])
    print(i, "value", round(d["value"], 4), "e2e", round(d["e2e"]["value"], 4), d["stages"]["ms"], d["stages"]["knn_ms_each_call"], d["e2e"]["seconds_each_step_rank0"])
PY
