"""Times the spectral initialisation on the BASELINE.json configs[1] graphs (texts 158,915 rows, images 31,783 rows,
16-D) and on a C4-shaped graph slice (1M rows, 2-D): the device-resident block eigensolver against the torch-assisted
round-1 form and torch.lobpcg (the reference's call, model.py:232).  CUDA events around whole solves."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multimodal-umap_b200")]
import torch
import bench
from umap_b200 import graph as G, spectral

def graph_of(x, k):
    idx, dist = G.knn_graph(x, x, k, True)
    col, w, _, _ = G.smooth_knn(idx, dist, "bisect")
    return G.fuzzy_union(col, w)

def residual(g, v):
    aval = spectral.normalized_adjacency(g)
    av = torch.zeros_like(v).index_add_(0, g.row.long(), aval[:, None] * v[g.col.long()])
    lam = (v * av).sum(0)
    return float((av - v * lam).norm(dim=0).max())

def timeit(label, fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{label:46s} {e0.elapsed_time(e1) / reps:9.3f} ms", flush=True)
    return out

data = bench.make_data(bench.WORKLOADS["c2"])
cases = [("texts 158915 rows, 16-D", graph_of(data["texts"].cuda(), 15), 16), ("images 31783 rows, 16-D", graph_of(data["images"].cuda(), 15), 16)]
c4 = bench.make_data(dict(bench.WORKLOADS["c4"], mods=[("blobs", 1000000, 128, "blobs")]))["blobs"].cuda()
cases.append(("C4-shaped 1M rows, k=30, 2-D", graph_of(c4, 30), 2))
del c4
os.environ["MMUMAP_SPECTRAL_DEBUG"] = "1"
for label, g, d in cases:
    torch.manual_seed(0)
    spectral.spectral_init(g, d, method="chebfsi")
os.environ["MMUMAP_SPECTRAL_DEBUG"] = "0"
for label, g, d in cases:
    for method in ("chebfsi", "chebfsi_torch") + (("lobpcg",) if g.n_rows < 200000 else ()):
        torch.manual_seed(0)
        v = timeit(f"{label}: {method}", lambda: spectral.spectral_init(g, d, method=method), reps=3)
        print(f"{'':46s} residual max |A v - theta v| = {residual(g, v):.2e}", flush=True)
