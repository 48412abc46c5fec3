#!/bin/bash
# 4-GPU sanity: parity of all multi-GPU paths at W=4, bench quick
mkdir -p gpurun_out
W=${W:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29681 scripts/check_multigpu.py > gpurun_out/o_parity$W.log 2>&1; echo "parity rc=$?"; grep -E "bit-exact|exchange|optimiser|epoch|OK|rror" gpurun_out/o_parity$W.log | tail -14 | cut -c1-300
timeout 900 $TR --master-port 29682 bench.py --gpus $W --steps 3 --warmup 2 --no-cpu-baseline --quality device > gpurun_out/o_bench$W.json 2> gpurun_out/o_bench$W.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/o_bench$W.json").read().strip().split("\n")[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["config"]["epoch_tail_kernel"], d["stages"]["ms"], "c3", d["stages"]["c3"]["fit_s"], d["stages"]["c3"]["stage_ms_rank0"], "q", d["quality"]["device_stream"]["similarity_test"])
PY
timeout 300 python -m pytest tests/test_gpu_graph.py -q --no-header -p no:cacheprovider -k union > gpurun_out/o_pytest_union.log 2>&1; echo "union tests rc=$?"; tail -3 gpurun_out/o_pytest_union.log | cut -c1-200
