#!/bin/bash
# last 1-GPU verification: everything the driver runs at round end
mkdir -p gpurun_out
rm -f gpurun_out/quality_tests.json
timeout 2400 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/u_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/u_pytest.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/u_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/u_smoke.log
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/u_bench.json").read().strip().split("\n")[-1])
print("C2 value", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], d["stages"]["ms"], d["stages"]["epoch_kernels_us_per_launch"])
print("roofline frac", d["roofline"]["frac"], "l2 frac", d["roofline"]["l2_roof"]["frac"], "knn frac", d["roofline_other"]["frac"], "clocks", d["clocks"])
print("c3", d["stages"]["c3"]["fit_s"], "transform", d["stages"]["transform_100k"]["seconds"])
PY
