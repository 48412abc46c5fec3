"""Is the C2 kNN stage's run-to-run spread (38 ms .. 83 ms for the same work) the power cap?  Times the texts search
call by call with NVML sampled every ~2 ms (SM clock, board power, event reasons), back to back and with idle gaps."""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multimodal-umap_b200")]
import torch
import pynvml as nv
import bench
from umap_b200 import graph as G

nv.nvmlInit(); h = nv.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], False
def loop():
    while not stop:
        samples.append((time.perf_counter(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetPowerUsage(h) / 1000.0,
                        nv.nvmlDeviceGetCurrentClocksEventReasons(h)))
        time.sleep(0.002)
th = threading.Thread(target=loop, daemon=True); th.start()

wl = bench.WORKLOADS["c2"]
data = bench.make_data(wl, seed=0)
from umap_b200 import knn_tc as KT
k = wl.get("k", 15)
for name, label, gap in (("texts", "back to back", 0.0), ("texts", "0.2 s idle between calls", 0.2), ("images", "back to back", 0.0),
                         ("images", "0.2 s idle between calls", 0.2), ("texts", "back to back again", 0.0)):
    x = data[name].cuda()
    print(f"{name}: {tuple(x.shape)} k={k}", flush=True)
    for it in range(8):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); G.knn_graph(x, x, k, True); e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
        ss = [s for s in samples if t0 <= s[0] <= t1]
        clk = [s[1] for s in ss] or [0]; pw = [s[2] for s in ss] or [0]
        rs = 0
        for s in ss: rs |= s[3]
        print(f"{label:26s} call {it}: {e0.elapsed_time(e1):6.1f} ms  sm MHz min/median {min(clk)}/{sorted(clk)[len(clk)//2]}  "
              f"power W max {max(pw):.0f}  reasons 0x{rs:x}  ({len(ss)} samples) fb {KT.last_stats.get('first_pass_uncertified')}/{KT.last_stats.get('fallback_rows')}", flush=True)
        time.sleep(gap)
stop = True
