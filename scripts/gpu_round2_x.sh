#!/bin/bash
# 2 GPUs: the whole GPU suite (includes tests/test_gpu_multi.py) and a short C2 bench -- the row-sharded upload path after
# the allocation change in impl/model.py::_Uploads
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/x_pytest2.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/x_pytest2.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus 2 --steps 5 --warmup 3 --no-c3 --quality off > gpurun_out/x_bench2.json 2> gpurun_out/x_bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/x_bench2.json").read().strip().split("\n")[-1])
print("2 GPUs value", round(d["value"], 4), "e2e", round(d["e2e"]["value"], 4), d["e2e"]["seconds_each_step_rank0"], d["stages"]["ms"], d["e2e"]["h2d_bytes_per_rank"])
PY
