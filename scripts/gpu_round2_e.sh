#!/bin/bash
# 2-GPU pass: graph-replayed epochs -- parity, then the bench at N=2 with and without graphs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29621 scripts/check_multigpu.py > gpurun_out/e_parity.log 2>&1; echo "parity rc=$?"; grep -E "exchange|epoch|OK|Error|error" gpurun_out/e_parity.log | tail -8
timeout 900 $TR --master-port 29622 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline --quality device > gpurun_out/e_bench2.json 2> gpurun_out/e_bench2.err; echo "bench2 rc=$?"
MMUMAP_EPOCH_GRAPH=0 timeout 900 $TR --master-port 29623 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline --quality off --no-c3 --no-transform > gpurun_out/e_bench2_nograph.json 2> gpurun_out/e_bench2_nograph.err; echo "bench2 nograph rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/e_bench2.json", "gpurun_out/e_bench2_nograph.json"):
    try:
        d = json.loads(open(f).read().strip().split("\n")[-1])
        print(f, "value", d["value"], "e2e", d["e2e"]["value"], d["stages"]["ms"], d["stages"]["epoch_kernels_us_per_launch"], d["stages"].get("transform_100k"), d.get("quality"))
    except Exception as ex:
        print(f, "unreadable", ex)
PY
tail -3 gpurun_out/e_bench2.err
