"""Why do single end-to-end fits (host buffers in, host result out) sometimes take 0.30-0.38 s instead of 0.22 s?
Per call: wall time, allocator counters (cudaMalloc calls, reserved/allocated bytes), objects only the cyclic GC frees,
and the pinned host->device bandwidth measured right before the call."""
import gc, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multimodal-umap_b200")]
import torch
import bench
from impl import util as util_mod

wl = bench.WORKLOADS["c2"]; OPT = bench.OPT
data = bench.make_data(wl, seed=0)
host = {k: v.pin_memory() for k, v in data.items()}
cfg = util_mod.Config(k_neighbors=wl["k"], out_dim=wl["out_dim"], min_dist=OPT["min_dist"], train_epochs=wl["epochs"],
                      num_rep=OPT["num_rep"], lr=OPT["lr"], alpha=OPT["alpha"], batch_size=OPT["batch_size"], test_epochs=120)
probe_src = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
probe_dst = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def h2d_gbs():
    torch.cuda.synchronize(); t0 = time.perf_counter()
    probe_dst.copy_(probe_src, non_blocking=True); torch.cuda.synchronize()
    return probe_src.numel() / (time.perf_counter() - t0) / 1e9
gc.collect()
for auto_gc in (True, False):
    (gc.enable if auto_gc else gc.disable)()
    for it in range(8):
        bw = h2d_gbs()
        st0 = torch.cuda.memory_stats()
        t0 = time.perf_counter()
        torch.manual_seed(1234)
        model = util_mod.train(host, cfg)
        out = [e.detach().cpu() for e in model.embeds]
        del model
        dt = time.perf_counter() - t0
        st1 = torch.cuda.memory_stats()
        print(f"gc {'auto' if auto_gc else 'off '} call {it}: {dt * 1e3:6.1f} ms  h2d before {bw:5.1f} GB/s  cudaMalloc calls {st1['num_device_alloc'] - st0['num_device_alloc']:3d} "
              f"cudaFree {st1['num_device_free'] - st0['num_device_free']:3d}  allocated {torch.cuda.memory_allocated() / 2**30:5.2f} GiB reserved {torch.cuda.memory_reserved() / 2**30:5.2f} GiB", flush=True)
    n = gc.collect()
    print(f"   gc.collect() found {n} unreachable objects; allocated after {torch.cuda.memory_allocated() / 2**30:5.2f} GiB", flush=True)
# what the per-fit cyclic garbage is made of
import collections
gc.enable(); gc.collect(); gc.set_debug(gc.DEBUG_SAVEALL)
torch.manual_seed(1234)
model = util_mod.train(host, cfg); out = [e.detach().cpu() for e in model.embeds]; del model, out
gc.collect()
print("cyclic garbage of one fit:", collections.Counter(type(o).__name__ for o in gc.garbage).most_common(12), flush=True)
for o in gc.garbage:
    if type(o).__name__ in ("function", "cell", "dict") and len(repr(o)) < 300:
        print("   ", type(o).__name__, repr(o)[:200])
gc.set_debug(0); gc.garbage.clear()
