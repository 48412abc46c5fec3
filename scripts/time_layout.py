"""Times optimiser epochs alone on C2-shaped stand-in graphs (random k-regular pattern, weights with
the measured mean) -- CUDA events, device sample stream."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multimodal-umap_b200")]
import torch
from umap_b200.graph import Graph
from umap_b200.layout import LayoutOptimizer

def standin(n, deg, wmean, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    col = torch.randint(0, n, (n, deg), generator=g, device="cuda", dtype=torch.int32).sort(dim=1).values
    w = (torch.rand((n, deg), generator=g, device="cuda") * 2 * wmean).clamp(max=1.0)
    return Graph.from_fixed_degree(col, w, n)

d = int(os.environ.get("DIM", "16"))
mods = [(158915, 26, 0.277), (31783, 25, 0.29)]
graphs = [standin(n, deg, wm, i) for i, (n, deg, wm) in enumerate(mods)]
embeds = [torch.randn((n, d), device="cuda") * 0.01 for n, _, _ in mods]
opt = LayoutOptimizer(embeds, graphs, 1.577, 0.8951, 8, 0.01, 1.0, 256, mode="fit", sample_stream="device", seed=1)
opt.run(20)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
E = 200
e0.record(); opt.run(E); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / E
kept = opt.kept_last_epoch()
from umap_b200 import profiler
profiler.enable(2)
opt.run(50)
for k, v in profiler.summarize(profiler.collect()).items():
    print(f"   {k:14s} {v['ms']/v['calls']*1e3:8.1f} us/launch x {v['calls']//50}/epoch")
profiler.enable(False)
print(f"{ms*1e3:.1f} us/epoch  kept={kept}  {kept*9/ms/1e6:.2f} G edge-updates/s", flush=True)
