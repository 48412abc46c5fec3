import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multimodal-umap_b200")]
import torch
from umap_b200 import knn_tc
from scripts.time_knn import data
n = int(os.environ.get("N", "1000000"))
x = data(n, 768, "bert")
for ms_ in (0, 2, 4, 8):
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        idx, dist, st, fb = knn_tc._call(x, x, 15, True, 0, None, True, ms_, 0)
        e1.record(); torch.cuda.synchronize()
    print(f"pairs={os.environ.get('MMUMAP_KNN_CTA_PAIRS','1')} n={n} min_splits={ms_}: {e0.elapsed_time(e1):.1f} ms  uncertified={st[0]}", flush=True)
