"""Times optimiser epochs on a C4-shaped problem (10M x 2-D, k=30: union degree ~51, mean weight ~0.167 -> ~85M kept
edges per epoch) with a stand-in graph (random k-regular pattern) -- CUDA events, device sample stream -- for the
tail-window settings of the force kernel (option sgd_window_mb).  Usage: time_layout_c4.py [n_rows] [window_mb ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multimodal-umap_b200")]
import torch
from umap_b200 import native, profiler
from umap_b200.graph import Graph
from umap_b200.layout import LayoutOptimizer

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
windows = [int(a) for a in sys.argv[2:]] or [0, 48]
d, deg, wmean = int(os.environ.get("DIM", "2")), 51, 0.167
epochs = int(os.environ.get("EPOCHS", "20"))
g = torch.Generator(device="cuda").manual_seed(0)
col = torch.empty((n, deg), dtype=torch.int32, device="cuda")
for lo in range(0, n, 1_000_000):
    hi = min(n, lo + 1_000_000)
    col[lo:hi] = torch.randint(0, n, (hi - lo, deg), generator=g, device="cuda", dtype=torch.int32).sort(dim=1).values
w = (torch.rand((n, deg), generator=g, device="cuda") * 2 * wmean).clamp(max=1.0)
graph = Graph.from_fixed_degree(col, w, n)
del col, w
embed = torch.randn((n, d), device="cuda") * 0.01
peak = 6550.0
# measured roofs of this access shape (csrc/roofs.cu): random 8-byte row gathers + reds on an L2-resident slice (one tail
# window) and on the whole table (DRAM sectors)
from umap_b200.native import check, lib, ptr, stream
for label, rows in (("one 48 MB window (L2)", 48 * 2 ** 20 // (2 * d * 4)), ("whole table (DRAM sectors)", n)):
    tab = torch.randn((rows, d), device="cuda"); acc = torch.zeros((rows, d), device="cuda"); sink = torch.zeros(1, device="cuda")
    touched = 400_000_000
    best = 1e9
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib().mmu_roof_random_rows(ptr(tab), ptr(acc), rows, d, touched, 5 + it, 1, 1, ptr(sink), stream()), "roof")
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"roof, random {d * 4}-byte rows, {label}: {2 * touched / best / 1e6:.1f} G row accesses/s "
          f"({2 * touched * d * 4 / best / 1e6:.0f} GB/s payload); an epoch's 18 tail accesses x 85.2M kept edges need "
          f"{18 * 85.2e6 / (2 * touched / best) :.2f} ms at that rate", flush=True)
    del tab, acc
for wmb in windows:
    native.set_option("sgd_window_mb", wmb)
    opt = LayoutOptimizer([embed], [graph], 1.577, 0.8951, 8, 0.01, 1.0, 256, mode="fit", sample_stream="device", seed=1)
    opt.run(3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); opt.run(epochs); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / epochs
    kept = opt.kept_last_epoch()
    nnz = graph.nnz
    nbytes = 12 * nnz + kept * 10 * d * 4 * 2 + 28 * n * d          # SURVEY.md 8(d) canonical bytes per epoch
    profiler.enable(2)
    opt.run(5)
    rows = profiler.summarize(profiler.collect())
    profiler.enable(0)
    parts = ", ".join(f"{k} {v['ms'] / 5:.2f} ms" for k, v in rows.items())
    print(f"n={n} d={d} sgd_window_mb={wmb} window_rows={opt.mods[0].window_rows} kernel={native.last_kernel('edge_forces')}: "
          f"{ms:.2f} ms/epoch, kept={kept}, {kept * 9 / ms / 1e6:.1f} G edge-updates/s, algorithmic {nbytes / 1e9:.1f} GB/epoch = "
          f"{nbytes / ms / 1e6:.0f} GB/s = {nbytes / ms / 1e6 / peak:.3f} of the HBM copy peak  [{parts}]", flush=True)
    del opt
    torch.cuda.empty_cache()
