"""Per-iteration diagnostics of the spectral solvers on C2-shaped graphs (GPU)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multimodal-umap_b200")]
import torch
from umap_b200 import graph as G, spectral
from scripts.time_knn import data  # noqa

def build(n, d, kind):
    x = data(n, d, kind)
    idx, dist = G.knn_graph(x, x, 15, True)
    col, w, _, _ = G.smooth_knn(idx, dist, "bisect")
    return G.fuzzy_union(col, w)

def main():
  for name, n, d, kind in [("texts", 158915, 768, "bert"), ("images", 31783, 4096, "vae")]:
      g = build(n, d, kind)
      for method in ("lobpcg", "chebfsi"):
          torch.manual_seed(0)
          spectral.spectral_init(g, 16, method=method)
          torch.cuda.synchronize(); t0 = time.perf_counter()
          os.environ["MMUMAP_SPECTRAL_DEBUG"] = "1"
          v = spectral.spectral_init(g, 16, method=method)
          os.environ["MMUMAP_SPECTRAL_DEBUG"] = "0"
          torch.cuda.synchronize(); dt = time.perf_counter() - t0
          aval = spectral.normalized_adjacency(g)
          av = G.spmm(g, v.contiguous(), aval)
          th = (v * av).sum(0)
          res = (av - v * th).norm(dim=0)
          print(f"{name} {method}: {dt*1e3:.1f} ms  theta[min,max]=({float(th.min()):.6f},{float(th.max()):.6f})  res max={float(res.max()):.2e}", flush=True)


if __name__ == "__main__":
    main()
