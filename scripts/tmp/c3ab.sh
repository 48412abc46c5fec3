M=sm__cycles_elapsed.avg.per_second,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,lts__t_sectors_srcunit_tex.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed
WARM=0 REPS=1 timeout -k 10 200 ncu --metrics $M --clock-control none -k regex:knn_tc_candidates --launch-count 1 --csv --log-file gpurun_out/c3_pair_metrics.csv python scripts/time_knn.py c3 > /dev/null 2>&1
WARM=0 REPS=1 MMUMAP_KNN_CTA_PAIRS=0 timeout -k 10 200 ncu --metrics $M --clock-control none -k regex:knn_tc_candidates --launch-count 1 --csv --log-file gpurun_out/c3_nopair_metrics.csv python scripts/time_knn.py c3 > /dev/null 2>&1
python - <<'PY'
import csv
for f in ["gpurun_out/c3_pair_metrics.csv","gpurun_out/c3_nopair_metrics.csv"]:
    print(f)
    rows=[l for l in open(f) if not l.startswith("==")]
    for r in csv.DictReader(rows):
        print("  ", r["Metric Name"], r["Metric Value"], r["Metric Unit"])
PY
