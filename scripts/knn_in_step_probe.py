"""The C2 kNN stage INSIDE consecutive fits (it starts right after the previous fit's 600 optimiser epochs have held the
board near its power limit): per fit, the stage time of each modality's search with the SM clock / power NVML reported
during it (2 ms sampling), then the same searches after a 0.3 s idle gap before each fit."""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multimodal-umap_b200")]
import torch
import pynvml as nv
import bench
from impl import util as util_mod
from umap_b200 import graph as G

nv.nvmlInit(); h = nv.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], False
def loop():
    while not stop:
        samples.append((time.perf_counter(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetPowerUsage(h) / 1000.0,
                        nv.nvmlDeviceGetCurrentClocksEventReasons(h)))
        time.sleep(0.002)
threading.Thread(target=loop, daemon=True).start()

wl = bench.WORKLOADS["c2"]
OPT = bench.OPT
data = {k: v.cuda() for k, v in bench.make_data(wl, seed=0).items()}
cfg = util_mod.Config(k_neighbors=wl["k"], out_dim=wl["out_dim"], min_dist=OPT["min_dist"], train_epochs=wl["epochs"],
                      num_rep=OPT["num_rep"], lr=OPT["lr"], alpha=OPT["alpha"], batch_size=OPT["batch_size"], test_epochs=120)
calls = []
inner = G.knn_graph
def wrapped(q, db, *a, **kw):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = inner(q, db, *a, **kw)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    calls.append((tuple(db.shape), t0, t1))
    return r
G.knn_graph = wrapped
from umap_b200 import profiler
keep = None
for label, gap, hold, prof, sync_wrap in (("consecutive fits", 0.0, False, 0, True), ("previous model kept alive", 0.0, True, 0, True),
                                          ("kept alive + stage events", 0.0, True, 1, True), ("same, no host sync around kNN", 0.0, True, 1, False),
                                          ("consecutive fits again", 0.0, False, 0, True)):
    G.knn_graph = wrapped if sync_wrap else inner
    for it in range(6):
        time.sleep(gap)
        calls.clear()
        profiler.enable(prof)
        torch.cuda.synchronize(); f0 = time.perf_counter()
        torch.manual_seed(1234); m = util_mod.train(data, cfg); torch.cuda.synchronize()
        f1 = time.perf_counter()
        if hold:
            keep = m
        else:
            keep = None
        del m
        if prof:
            st = profiler.summarize(profiler.collect())
            label_x = "  events: " + ", ".join(f"{k}={v['ms']:.1f}" for k, v in st.items())
        else:
            label_x = ""
        profiler.enable(0)
        parts = []
        for shape, t0, t1 in calls:
            ss = [s for s in samples if t0 <= s[0] <= t1]
            clk = [s[1] for s in ss] or [0]; pw = [s[2] for s in ss] or [0]
            parts.append(f"{shape[0]}x{shape[1]}: {(t1 - t0) * 1e3:5.1f} ms (sm MHz {min(clk)}..{max(clk)}, {max(pw):.0f} W)")
        allp = [s[2] for s in samples if f0 <= s[0] <= f1]
        print(f"{label:28s} fit {it}: {(f1 - f0) * 1e3:6.1f} ms   kNN " + "  ".join(parts) + f"   fit power max {max(allp):.0f} W" + label_x, flush=True)
stop = True
