"""Wall-clock breakdown of spectral_chebfsi on the C2 texts graph (sync after each section)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multimodal-umap_b200")]
import torch
from umap_b200 import graph as G, spectral
from scripts.spectral_diag import build  # noqa  (runs its main too; cheap enough)
from scripts.time_knn import data
g = build(158915, 768, "bert")
def T(label, fn, n=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): out = fn()
    torch.cuda.synchronize(); print(f"{label:28s} {(time.perf_counter()-t0)/n*1e3:8.3f} ms"); return out
aval = T("normalized_adjacency", lambda: spectral.normalized_adjacency(g))
n, b = g.n_rows, 32
n_pad = -(-n // 256) * 256
x = torch.zeros((n_pad, b), device="cuda"); x[:n] = torch.randn((n, b), device="cuda")
T("randn init", lambda: torch.randn((n, b), device="cuda"))
T("gram", lambda: spectral._gram(x, x))
gm = spectral._gram(x, x); gm = 0.5 * (gm + gm.T)
T("eigh 32x32 (gpu)", lambda: torch.linalg.eigh(gm))
T("eigh 32x32 (cpu roundtrip)", lambda: [t.cuda() for t in torch.linalg.eigh(gm.cpu())])
T("orthonormalise", lambda: spectral._orthonormalise(x))
y = torch.zeros_like(x)
T("spmm_axpby b=32", lambda: G.spmm_axpby(g, aval, x, 1.0, 0.0, None, 0.0, out=y))
v = torch.randn((b, b), device="cuda")
T("x @ v", lambda: x @ v)
T("residual norms", lambda: (y[:, :17] - x[:, :17] * v[0, :17]).norm(dim=0).max())
T("float(sync)", lambda: float(gm[0, 0]))
T("full chebfsi", lambda: spectral.spectral_chebfsi(g, 16), n=3)
