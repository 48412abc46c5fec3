#!/bin/bash
# final 1-GPU pass: tests, smoke, bench (engine + reference arm), launch list, kNN ncu, C4
mkdir -p gpurun_out
rm -f gpurun_out/quality_tests.json
timeout 2400 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/m_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/m_pytest.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/m_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/m_smoke.log
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err; echo "bench rc=$?"
timeout 1500 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/m_bench_reference.json 2> gpurun_out/m_bench_reference.err; echo "reference arm rc=$?"; tail -c 1500 gpurun_out/m_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/m_launches.csv \
    python bench.py --steps 1 --warmup 1 --quick --epochs 12 --no-cpu-baseline --no-transform > gpurun_out/m_ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_tc_candidates -s 2 -c 2 -f -o gpurun_out/m_knn_c2 \
    python scripts/time_knn.py > gpurun_out/m_ncu_knn_c2.log 2>&1; echo "knn c2 ncu rc=$?"
MMUMAP_BENCH_DEBUG=1 timeout 1500 python bench.py --workload c4 --steps 1 --warmup 1 --quick --no-cpu-baseline > gpurun_out/m_bench_c4.json 2> gpurun_out/m_bench_c4.err; echo "bench c4 rc=$?"; grep "stages ms" gpurun_out/m_bench_c4.err | tail -1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/m_bench.json").read().strip().split("\n")[-1])
print("C2 value", d["value"], "e2e", d["e2e"], "launches", d["gpu_launches"], d["stages"]["ms"], d["stages"]["epoch_kernels_us_per_launch"])
print("roofline", {k: d["roofline"][k] for k in ("kernel", "achieved", "frac", "traffic", "effective_bound", "share_of_step", "ms_per_launch")}, d["roofline"]["l2_roof"]["frac"])
print("quality", d["quality"]); print("c3", d["stages"]["c3"]); print("transform", d["stages"]["transform_100k"]); print("cpu_baseline", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"])
c = json.loads(open("gpurun_out/m_bench_c4.json").read().strip().split("\n")[-1])
print("C4 value", c["value"], c["stages"]["ms"])
PY
