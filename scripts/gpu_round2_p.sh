#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
MMUMAP_BENCH_DEBUG=1 timeout 1500 $TR --master-port 29691 bench.py --gpus 2 --workload c4 --steps 1 --warmup 1 --quick --no-cpu-baseline > gpurun_out/p_bench_c4_2.json 2> gpurun_out/p_bench_c4_2.err
echo "bench c4 x2 rc=$?"; grep "rank 0\] stages ms" gpurun_out/p_bench_c4_2.err | tail -1
python - <<'PY'
import json
c = json.loads(open("gpurun_out/p_bench_c4_2.json").read().strip().split("\n")[-1])
print("C4 x2 value", c["value"], c["stages"]["ms"], c["config"]["epoch_tail_kernel"])
PY
