#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/h_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/h_pytest.log | cut -c1-300
MMUMAP_BENCH_DEBUG=1 timeout 1500 python bench.py --workload c4 --steps 1 --warmup 1 --quick --no-cpu-baseline > gpurun_out/h_bench_c4.json 2> gpurun_out/h_bench_c4.err
echo "bench c4 rc=$?"; grep "stages ms" gpurun_out/h_bench_c4.err | tail -2; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/h_bench_c4.json").read().strip().split("\n")[-1])
    print("C4 value", d["value"], d["stages"]["ms"], "knn_tflops", d["stages"]["knn_tflops"], "sgd frac/gpu", d["stages"]["sgd_hbm_frac_per_gpu"], "kept", d["stages"]["kept_edges_last_epoch"])
    print(d["roofline"]["kernel"], d["roofline"]["ms_per_launch"], d["roofline_other"]["kernel"])
except Exception as ex:
    print("unreadable", ex)
PY
tail -3 gpurun_out/h_bench_c4.err | cut -c1-400
