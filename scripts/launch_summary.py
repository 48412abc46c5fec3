"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (name + grid):
    python scripts/launch_summary.py gpurun_out/launches.csv"""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", "")) / 1e3
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:72] + " grid=" + row["Grid Size"]
    agg[name][0] += 1; agg[name][1] += v; tot += v
print(f"# {sys.argv[1]}: {sum(c for c, _ in agg.values())} launches, {tot/1e3:.2f} ms of kernel time (cold-cache, serialised by ncu: compare shares)")
print(f"{'total us':>12} {'launches':>8} {'us/launch':>11} {'share':>6}  kernel")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:12.1f} {c:8d} {t/c:11.1f} {100*t/tot:5.1f}%  {k}")
