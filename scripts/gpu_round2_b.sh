#!/bin/bash
# second GPU pass: full test suite, C4 window sweep + roofs, ncu captures
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -15 gpurun_out/b_pytest.log
timeout 400 python scripts/time_layout_c4.py 10000000 48 64 80 96 > gpurun_out/b_c4_layout.log 2>&1
tail -6 gpurun_out/b_c4_layout.log
# launch list of a short fit (12 epochs)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/b_launches.csv \
    python bench.py --steps 1 --warmup 1 --quick --epochs 12 --no-cpu-baseline --no-transform > gpurun_out/b_ncu_launch.log 2>&1
echo "launch list rc=$?"
# --set full: force kernel on C2-shaped tables (texts, images), then the windowed form on 10M x 2-D
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edge_forces_staged -s 40 -c 2 -f -o gpurun_out/b_forces_c2 \
    python scripts/time_layout.py > gpurun_out/b_ncu_forces_c2.log 2>&1
echo "forces c2 rc=$?"
EPOCHS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:edge_forces_staged -s 12 -c 4 -f -o gpurun_out/b_forces_c4 \
    python scripts/time_layout_c4.py 10000000 48 > gpurun_out/b_ncu_forces_c4.log 2>&1
echo "forces c4 rc=$?"
# --set full: short-row kNN candidates (1M x 128, k=30: split-fp16 level)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_tc_candidates -s 1 -c 1 -f -o gpurun_out/b_knn_smallD \
    python scripts/time_knn_c4.py > gpurun_out/b_ncu_knn_smallD.log 2>&1
echo "knn smallD rc=$?"
ls -la gpurun_out/*.ncu-rep
