"""Turns the `ncu --set full` captures of a round into the tracked files bench.py and the judge read:

    python scripts/ncu_facts.py <round-tag> <site>=<capture.ncu-rep>[:<label,label,...>] ...

e.g.  python scripts/ncu_facts.py r02 edge_forces=gpurun_out/b_forces_c2.ncu-rep:texts,images knn_candidates=...

For every capture: a text summary (scripts/ncu_summary.py's metric list) under profiles/<tag>_<site>_ncu_full.txt, and one
entry in profiles/<tag>_ncu_facts.json with, per launch, the kernel name, duration, DRAM bytes (read + write), L2 / L1TEX /
tensor-pipe utilisation -- bench.py takes its `traffic` and ncu figures from that file, never from constants.  Runs here
(no GPU needed: `ncu -i` only reads the report)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FACTS = {
    "gpu__time_duration.sum": "duration_ns",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_pct",
    "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed": "l1tex2xbar_req_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "lts__t_sectors.sum": "l2_sectors",
    "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed": "l2_atomic_input_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__cycles_elapsed.avg.per_second": "sm_clock_hz",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
}
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9,
              "hz": 1.0, "Khz": 1e3, "Mhz": 1e6, "Ghz": 1e9}


def parse(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    launches = []
    for row in rows[2:]:
        rec = {"kernel": row[hdr.index("Kernel Name")]}
        for i, k in enumerate(hdr):
            if k in FACTS and row[i] not in ("", "n/a"):
                try:
                    v = float(row[i].replace(",", ""))
                except ValueError:
                    continue
                rec[FACTS[k]] = v * UNIT_SCALE.get(units[i], 1.0)
        launches.append(rec)
    return launches


def main():
    tag = sys.argv[1]
    facts_path = os.path.join(ROOT, "profiles", f"{tag}_ncu_facts.json")
    facts = json.load(open(facts_path)) if os.path.exists(facts_path) else {}
    for spec in sys.argv[2:]:
        site, rest = spec.split("=", 1)
        rep, _, labels = rest.partition(":")
        labels = labels.split(",") if labels else []
        launches = parse(rep)
        txt = os.path.join(ROOT, "profiles", f"{tag}_{site}_ncu_full.txt")
        summ = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep, "l1tex2xbar", "lts__t_sectors.sum",
                               "sm__cycles_elapsed.avg.per_second"], capture_output=True, text=True).stdout
        with open(txt, "w") as f:
            f.write(f"# ncu --set full --clock-control none, capture {os.path.basename(rep)} (launch labels: {labels or 'in order'})\n")
            f.write(summ)
        per = {}
        for i, l in enumerate(launches):
            lab = labels[i] if i < len(labels) else f"launch{i}"
            l["dram_bytes"] = l.get("dram_read_bytes", 0.0) + l.get("dram_write_bytes", 0.0)
            per[lab] = l
        entry = {"source": os.path.relpath(txt, ROOT), "kernel": launches[0]["kernel"] if launches else None, "launches": per,
                 # `traffic` of bench.py's roofline object: DRAM read + write bytes per launch, averaged over the captured launches
                 "dram_bytes_per_launch": sum(l["dram_bytes"] for l in launches) / max(len(launches), 1)}
        facts[site] = entry
        print(f"{site}: {len(launches)} launches -> {txt}")
    json.dump(facts, open(facts_path, "w"), indent=1, sort_keys=True)
    print("wrote", facts_path)


if __name__ == "__main__":
    main()
