"""Times one rank's share of the row-sharded kNN (query rows [lo,hi) against the full database)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multimodal-umap_b200")]
import torch
from umap_b200 import knn_tc, dist as D
from scripts.time_knn import data
W = int(os.environ.get("W", "4"))
for name, n, d, kind in [("texts", 158915, 768, "bert"), ("images", 31783, 4096, "vae")]:
    x = data(n, d, kind)
    for r in (0, W - 1):
        lo, hi = D.row_block(n, r, W)
        q = x[lo:hi]
        for _ in range(2):
            knn_tc.knn_tc(q, x, 15, True, query_base=lo)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            knn_tc.knn_tc(q, x, 15, True, query_base=lo)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"{name} rank {r}/{W} rows {hi-lo}: {ms:.2f} ms  {2.0*(hi-lo)*n*d/ms/1e9:.1f} TFLOP/s  {knn_tc.last_stats}", flush=True)
