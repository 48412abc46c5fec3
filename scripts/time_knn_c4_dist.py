"""C4-shaped kNN (N x 128 blobs, k=30) through graph.knn_graph under torch.distributed.run: the cluster-pruned search with
its query blocks sharded over the ranks, checked against the single-rank result on a row sample."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multimodal-umap_b200")]
import torch
import torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from umap_b200 import dist as D, graph as G, knn_tc
rank, world = dist.get_rank(), dist.get_world_size()
n, d, k = int(os.environ.get("N", "4000000")), 128, 30
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.empty((n, d), device="cuda")
cent = 5.0 * torch.randn((1000, d), generator=g, device="cuda")
for lo in range(0, n, 1000000):
    hi = min(n, lo + 1000000)
    x[lo:hi] = cent[torch.arange(lo, hi, device="cuda") % 1000] + torch.randn((hi - lo, d), generator=g, device="cuda")
for it in range(2):
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); idx, dd = G.knn_graph(x, x, k, True); e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"{world} GPUs {n}x{d} k={k}: {e0.elapsed_time(e1):.1f} ms  {knn_tc.last_stats}", flush=True)
# parity on a row sample against the exhaustive kernel
rows = torch.arange(0, n, n // 2000, device="cuda", dtype=torch.int32)[:2000]
ei, ed = torch.full_like(idx, -7), torch.full_like(dd, -7.0)
G.knn_exact_simt(x, x, k, True, out=(ei, ed), rows=rows)
r = rows.long()
ok = bool(torch.equal(idx[r], ei[r])) and bool(torch.equal(dd[r].view(torch.int32), ed[r].view(torch.int32)))
print(f"[rank {rank}] sample of 2000 rows bit-exact vs the exhaustive kernel: {ok}", flush=True)
assert ok
dist.destroy_process_group()
