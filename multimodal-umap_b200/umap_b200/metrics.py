"""Batched GPU equivalents of the reference's acceptance metrics.

/root/reference/impl/validation.py:57-84 (knn_test) loops over the queries in Python with a host
sync per row (`idx in knns`), which dominates BASELINE.json configs[4] at 100k queries; here the
retrieval runs through the engine's exact kNN kernel in query mode.  validation.py itself keeps
running unchanged on top of the engine and is the parity check at small Q
(tests/test_gpu_e2e.py::test_batched_metrics_equal_the_row_loop).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import graph as G


def retrieval_accuracy(src_embed: torch.Tensor, dst_embed: torch.Tensor, k: int) -> float:
    """Fraction of rows i whose counterpart i is among the k nearest rows of the other table, both
    directions averaged (the inner loop of validation.py:66-78, ties broken towards the smaller
    index where torch.topk leaves them unspecified)."""
    n = src_embed.shape[0]
    a = src_embed.detach().to("cuda", torch.float32).contiguous()
    b = dst_embed.detach().to("cuda", torch.float32).contiguous()
    own = torch.arange(n, device=a.device, dtype=torch.int32)[:, None]
    fwd, _ = G.knn_graph(a, b, k, exclude_self=False)
    bwd, _ = G.knn_graph(b, a, k, exclude_self=False)
    correct = (fwd == own).any(dim=1).sum() + (bwd == own).any(dim=1).sum()
    return float(correct.item()) / (2 * n)


def retrieval_accuracies(src_embed: torch.Tensor, dst_embed: torch.Tensor, ks) -> dict:
    """retrieval_accuracy for several k from ONE search at max(ks): the engine's rows are sorted by
    (distance, index), so the k nearest are the first k columns."""
    n = src_embed.shape[0]
    a = src_embed.detach().to("cuda", torch.float32).contiguous()
    b = dst_embed.detach().to("cuda", torch.float32).contiguous()
    own = torch.arange(n, device=a.device, dtype=torch.int32)[:, None]
    kmax = max(ks)
    fwd, _ = G.knn_graph(a, b, kmax, exclude_self=False)
    bwd, _ = G.knn_graph(b, a, kmax, exclude_self=False)
    return {k: float(((fwd[:, :k] == own).any(dim=1).sum() + (bwd[:, :k] == own).any(dim=1).sum()).item()) / (2 * n)
            for k in ks}


def knn_test_multi(model, embed_fn, data: dict, cfg, ks=(1, 5)) -> dict:
    """knn_test of validation.py:40-84 for several k with one transform per modality pair."""
    mats = [data[key] for key in data]
    accs = {k: [] for k in ks}
    for src in range(len(mats)):
        for dst in range(src + 1, len(mats)):
            e = embed_fn(model, [mats[src], mats[dst]], [src, dst], cfg)
            for k, v in retrieval_accuracies(e[0], e[1], ks).items():
                accs[k].append(v)
    return {k: float(torch.tensor(v).mean().item()) for k, v in accs.items()}


def knn_test(model, embed_fn, data: dict, cfg, k: int = 5) -> float:
    """knn_test of validation.py:40-84 with the batched retrieval."""
    mats = [data[key] for key in data]
    accs = []
    for src in range(len(mats)):
        for dst in range(src + 1, len(mats)):
            e = embed_fn(model, [mats[src], mats[dst]], [src, dst], cfg)
            accs.append(retrieval_accuracy(e[0], e[1], k))
    return float(torch.tensor(accs).mean().item())


def similarity_test(model, embed_fn, data: dict, cfg) -> float:
    """similarity_test of validation.py:7-38."""
    mats = [data[key] for key in data]
    embeds = [F.normalize(e.detach(), p=2, dim=1) for e in embed_fn(model, mats, list(range(len(mats))), cfg)]
    sims = [(embeds[i] * embeds[j]).sum(dim=1) for i in range(len(mats)) for j in range(i + 1, len(mats))]
    return float(torch.stack(sims, dim=1).mean(dim=1).mean().item())
