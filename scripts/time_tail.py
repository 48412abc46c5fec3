"""Times the multi-GPU epoch tail ALONE (no force kernels in front of it: the ranks arrive together) on buffers of the
BASELINE.json configs[1] size (158,915 + 31,783 rows x 16-D = 12.2 MB), for the exchange forms and grid sizes.
Launch with torch.distributed.run, one rank per GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multimodal-umap_b200")]
import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from umap_b200 import native
from umap_b200.graph import Graph
from umap_b200.layout import LayoutOptimizer
rank, world = dist.get_rank(), dist.get_world_size()

def tiny_graph(n):
    col = torch.arange(n, dtype=torch.int32, device="cuda").roll(1)[:, None].contiguous()
    return Graph.from_fixed_degree(col, torch.full((n, 1), 1e-6, device="cuda"), n)

rows = [158915, 31783]
embeds = [torch.randn((n, 16), device="cuda") * 0.01 for n in rows]
graphs = [tiny_graph(n) for n in rows]
for label, env, per_sm in (("push", {"MMUMAP_PEER_TAIL": "push"}, 1), ("push, no gradient push [timing only]", {"MMUMAP_PEER_TAIL": "push", "_skip": "1"}, 1),
                           ("push, no shard step [timing only]", {"MMUMAP_PEER_TAIL": "push", "_skip": "2"}, 1),
                           ("push, barriers only [timing only]", {"MMUMAP_PEER_TAIL": "push", "_skip": "3"}, 1), ("fused multimem 1/SM", {"MMUMAP_PEER_TAIL": "fused"}, 1), ("fused peer loads 1/SM", {"MMUMAP_PEER_TAIL": "fused", "MMUMAP_PEER_MULTIMEM": "0"}, 1),
                           ("legacy 5 launches", {"MMUMAP_PEER_TAIL": "legacy"}, 1)):
    os.environ.pop("MMUMAP_PEER_TAIL", None)
    env = dict(env)
    native.set_option("tail_skip_mask", int(env.pop("_skip", "0")))
    os.environ.update(env)
    native.set_option("tail_blocks_per_sm", per_sm)
    import umap_b200.layout as LY
    if "MMUMAP_PEER_MULTIMEM" in env:
        LY._PEER_CACHE.clear()
    opt = LayoutOptimizer(embeds, graphs, 1.577, 0.8951, 8, 0.01, 1.0, 256, mode="fit", sample_stream="device", seed=1)
    for _ in range(20):
        opt._adam_tail()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(300):
        opt._adam_tail()
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"{world} GPUs, {label:40s} {native.last_kernel('epoch_tail') or 'barrier + adam_peer + barrier + clear':42s} {e0.elapsed_time(e1) / 300 * 1e3:7.1f} us per tail", flush=True)
    dist.barrier()
    os.environ.pop("MMUMAP_PEER_MULTIMEM", None)
    if "MMUMAP_PEER_MULTIMEM" in env:
        LY._PEER_CACHE.clear()
dist.destroy_process_group()
