"""Times the invert-mode force kernel (mmu_invert_forces; /root/reference/impl/model.py:336-362) at the crossmodal shape:
Q = 100,000 reconstructed rows in the 4,096-D image-latent space, k = 15 neighbours, 8 negatives, against 31,783 fitted
rows -- CUDA events, device sample stream, a random stand-in graph."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multimodal-umap_b200")]
import torch
from umap_b200 import profiler
from umap_b200.graph import Graph
from umap_b200.layout import LayoutOptimizer
q, n, dim, k = int(os.environ.get("Q", "100000")), 31783, 4096, 15
g = torch.Generator(device="cuda").manual_seed(0)
col = torch.randint(0, n, (q, k), generator=g, device="cuda", dtype=torch.int32).sort(dim=1).values
w = torch.rand((q, k), generator=g, device="cuda") * 0.6
graph = Graph.from_fixed_degree(col, w, n)
data = torch.randn((n, dim), generator=g, device="cuda") * 4.0
x0 = torch.randn((q, dim), generator=g, device="cuda") * 4.0
sigma = torch.rand(n, generator=g, device="cuda") + 0.5
rho = torch.rand(n, generator=g, device="cuda") * 50
opt = LayoutOptimizer([x0], [graph], 1.577, 0.8951, 8, 0.01, 1.0, 256, mode="invert", refs=[data], sigmas=[sigma], rhos=[rho],
                      sample_stream="device", seed=1)
opt.run(3)
profiler.enable(2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); opt.run(10); e1.record(); torch.cuda.synchronize()
kept = opt.kept_last_epoch()
ms = e0.elapsed_time(e1) / 10
nbytes = kept * (1 + 9 + 1) * dim * 4          # per kept edge: x row read + 9 data rows (phase 1; phase 2 re-reads them from L2) + gradient row
print(f"invert epoch Q={q} D={dim}: {ms:.2f} ms, kept {kept}, {kept * 9 / ms / 1e6:.2f} G pair-updates/s, "
      f"{nbytes / ms / 1e6:.0f} GB/s of row traffic ({nbytes / ms / 1e6 / 6550:.2f} of the HBM copy peak)", flush=True)
