"""C4-shaped kNN at 1/10 scale: 1M x 128 blobs (1000 clusters), k=30."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multimodal-umap_b200")]
import torch
from umap_b200 import knn_tc
n, d, k = int(os.environ.get("N", "1000000")), 128, 30
g = torch.Generator(device="cuda").manual_seed(0)
x = (5.0 * torch.randn((1000, d), generator=g, device="cuda")[torch.arange(n, device="cuda") % 1000]
     + torch.randn((n, d), generator=g, device="cuda")).contiguous()
for _ in range(2):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); knn_tc.knn_tc(x, x, k, True); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{n}x{d} k={k}: {ms:.1f} ms  {2.0*n*n*d/ms/1e9:.1f} TFLOP/s (algorithmic 2QND)  {knn_tc.last_stats}", flush=True)
from umap_b200 import knn_pruned
print("contrast", knn_pruned.contrast(x, k), "pruned stats", knn_pruned.last_stats, flush=True)
