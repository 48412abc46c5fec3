import os, sys, ctypes
os.environ["MMUMAP_KNN_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multimodal-umap_b200")]
import numpy as np, torch
from umap_b200 import knn_tc, native
from scripts.time_knn import data
n = int(os.environ.get("N", "1000000"))
x = data(n, 768, "bert")
knn_tc._call(x, x, 15, True, 0, None, True, 0, 0)
torch.cuda.synchronize()
ncta = (n + 127) // 128
ncta += ncta & 1
buf = np.zeros(3 * ncta, dtype=np.uint64)
lib = native.lib()
lib.mmu_debug_knn_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
rc = lib.mmu_debug_knn_trace(buf.ctypes.data, ncta)
t = buf.reshape(-1, 3).astype(np.int64)
t0 = t[:, 1].min()
start = (t[:, 1] - t0) / 1e3; end = (t[:, 2] - t0) / 1e3; dur = end - start
tag = "pairs=" + os.environ.get("MMUMAP_KNN_CTA_PAIRS", "1")
print(tag, "rc", rc, "ctas", ncta, "kernel span us", end.max())
print(tag, "duration us: min %.0f p5 %.0f median %.0f p95 %.0f max %.0f" % (dur.min(), *np.percentile(dur, [5, 50, 95]), dur.max()))
order = np.argsort(start)
print(tag, "first-wave starts us (sorted, every 10th of first 148):", np.round(start[order][:148:10], 1))
print(tag, "first-wave durations us:", np.round(dur[order][:148:10], 0))
# phase spread: at the time the k-th CTA (by start) begins, what fraction of its pass has each running CTA done?
for k in (200, 2000, 6000):
    ts = start[order][k]
    running = (start <= ts) & (end > ts)
    phase = (ts - start[running]) / dur[running]
    print(tag, f"at t={ts:.0f} us: {running.sum()} CTAs running, phase min {phase.min():.3f} p25 {np.percentile(phase,25):.3f} median {np.median(phase):.3f} p75 {np.percentile(phase,75):.3f} max {phase.max():.3f}")
sm = t[:, 0]
per_sm = np.array([dur[sm == s].mean() for s in np.unique(sm)])
print(tag, "mean duration per SM: min %.0f max %.0f; SMs used %d" % (per_sm.min(), per_sm.max(), len(per_sm)))
np.save(os.path.join(ROOT, "gpurun_out", f"trace_{tag.replace('=','')}.npy"), t)
