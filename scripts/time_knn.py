"""Times the kNN stage alone on BASELINE.json configs[1] shapes (CUDA events, inputs resident)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multimodal-umap_b200")]
import torch
from umap_b200 import graph as G, knn_tc

def data(n, d, kind, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    cl = torch.arange(n, device="cuda") % 64
    if kind == "bert":
        return torch.tanh(torch.randn((64, d), generator=g, device="cuda")[cl] + 0.5 * torch.randn((n, d), generator=g, device="cuda")).contiguous()
    return (2.0 * torch.randn((64, d), generator=g, device="cuda")[cl] + 4.0 * torch.randn((n, d), generator=g, device="cuda")).contiguous()

def main():
    shapes = [("texts", 158915, 768, "bert"), ("images", 31783, 4096, "vae")]
    if len(sys.argv) > 1 and sys.argv[1] == "small":
        shapes = [("texts", 40000, 768, "bert"), ("images", 12000, 4096, "vae")]
    if len(sys.argv) > 1 and sys.argv[1] == "c3":
        shapes = [("c3", 1000000, 768, "bert")]
    reps = int(os.environ.get("REPS", "3"))
    for name, n, d, kind in shapes:
        x = data(n, d, kind)
        for _ in range(int(os.environ.get("WARM", "2"))):
            G.knn_graph(x, x, 15, True, method="tc")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            G.knn_graph(x, x, 15, True, method="tc")
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{name}: {ms:.2f} ms  {2.0*n*n*d/ms/1e9:.1f} TFLOP/s  stats={knn_tc.last_stats}", flush=True)



if __name__ == "__main__":
    main()
