#!/bin/bash
# final 8-GPU pass: parity of every multi-GPU path at W=8, C2 bench, C4 fit
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29701 scripts/check_multigpu.py > gpurun_out/q_parity8.log 2>&1; echo "parity rc=$?"; grep -E "rank 0.*(bit-exact|exchange|optimiser|epoch)|OK|rror" gpurun_out/q_parity8.log | tail -8 | cut -c1-300
timeout 900 $TR --master-port 29702 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --quality device > gpurun_out/q_bench8.json 2> gpurun_out/q_bench8.err; echo "bench8 rc=$?"
MMUMAP_BENCH_DEBUG=1 timeout 1500 $TR --master-port 29703 bench.py --gpus 8 --workload c4 --steps 1 --warmup 1 --quick --no-cpu-baseline > gpurun_out/q_bench_c4_8.json 2> gpurun_out/q_bench_c4_8.err
echo "bench c4 x8 rc=$?"; grep "rank 0\] stages ms" gpurun_out/q_bench_c4_8.err | tail -1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/q_bench8.json").read().strip().split("\n")[-1])
print("C2 x8 value", d["value"], "e2e", d["e2e"], d["config"]["epoch_tail_kernel"], d["stages"]["ms"], d["stages"]["epoch_kernels_us_per_launch"], d["stages"]["transform_100k"], "c3", d["stages"]["c3"], "quality", d["quality"])
c = json.loads(open("gpurun_out/q_bench_c4_8.json").read().strip().split("\n")[-1])
print("C4 x8 value", c["value"], c["stages"]["ms"], c["config"]["epoch_tail_kernel"])
PY
