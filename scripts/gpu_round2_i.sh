#!/bin/bash
# 8-GPU pass: epoch-tail forms alone, the C2 bench (default and peer-load tail), the C4 fit
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29641 scripts/time_tail.py > gpurun_out/i_tail8.log 2>&1; echo "tail rc=$?"; grep "us per tail\|Error\|error" gpurun_out/i_tail8.log | tail -12
timeout 900 $TR --master-port 29642 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --quality device > gpurun_out/i_bench8.json 2> gpurun_out/i_bench8.err; echo "bench8 rc=$?"
MMUMAP_PEER_MULTIMEM=0 MMUMAP_TAIL_BLOCKS_PER_SM=4 timeout 900 $TR --master-port 29643 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --quality off --no-c3 --no-transform > gpurun_out/i_bench8_peerloads.json 2> gpurun_out/i_bench8_peerloads.err; echo "bench8 peer loads rc=$?"
MMUMAP_EPOCH_GRAPH=0 timeout 900 $TR --master-port 29644 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --quality off --no-c3 --no-transform > gpurun_out/i_bench8_nograph.json 2> gpurun_out/i_bench8_nograph.err; echo "bench8 nograph rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/i_bench8.json", "gpurun_out/i_bench8_peerloads.json", "gpurun_out/i_bench8_nograph.json"):
    try:
        d = json.loads(open(f).read().strip().split("\n")[-1])
        print(f, "value", d["value"], "e2e", d["e2e"], d["config"]["epoch_tail_kernel"], d["stages"]["ms"], d["stages"]["epoch_kernels_us_per_launch"], d["stages"].get("transform_100k"), "c3", d["stages"].get("c3"))
    except Exception as ex:
        print(f, "unreadable", ex)
PY
MMUMAP_BENCH_DEBUG=1 timeout 1500 $TR --master-port 29645 bench.py --gpus 8 --workload c4 --steps 1 --warmup 1 --quick --no-cpu-baseline > gpurun_out/i_bench_c4_8.json 2> gpurun_out/i_bench_c4_8.err
echo "bench c4 x8 rc=$?"; grep "rank 0\] stages ms" gpurun_out/i_bench_c4_8.err | tail -1; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/i_bench_c4_8.json").read().strip().split("\n")[-1])
    print("C4 x8 value", d["value"], d["stages"]["ms"], "sgd frac/gpu", d["stages"]["sgd_hbm_frac_per_gpu"])
except Exception as ex:
    print("unreadable", ex)
PY
