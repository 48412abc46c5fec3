"""Tensor-core exact kNN (K1/K2, include/mmumap.h: mmu_knn_tc).

Role of /root/reference/impl/model.py:81-195 (candidate search + per-row top-k).  The tcgen05
kernel generates candidates, every row is certified and rescored with the canonical fp32
distance; rows that cannot be certified are finished by the exhaustive fp32 kernel, so the
result equals mmu_knn_exact_f32 bit for bit.
"""
from __future__ import annotations

import torch

from . import native
from .native import check, lib, ptr, stream

last_stats: dict = {}


def knn_tc(query: torch.Tensor, db: torch.Tensor, k: int, exclude_self: bool, query_base: int = 0):
    """Returns (idx int32 [Q,k] sorted by (dist, idx), dist float32 [Q,k])."""
    native.require_cuda()
    if k > native.KNN_TC_MAX_K:
        raise ValueError(f"tensor-core kNN supports k <= {native.KNN_TC_MAX_K}, got {k}")
    same = query is db or (query.data_ptr() == db.data_ptr() and query.shape == db.shape)
    db = db.detach().to("cuda", torch.float32).contiguous()
    query = db if same else query.detach().to("cuda", torch.float32).contiguous()
    same = same and query_base == 0
    q, d = query.shape
    n = db.shape[0]
    dev = db.device
    ws_bytes = lib().mmu_knn_tc_workspace_bytes(q, n, d, int(same))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    idx = torch.empty((q, k), dtype=torch.int32, device=dev)
    dist = torch.empty((q, k), dtype=torch.float32, device=dev)
    stats = torch.empty(4, dtype=torch.int32, device=dev)
    fallback = torch.empty(max(q, 1), dtype=torch.int32, device=dev)
    check(lib().mmu_knn_tc(ptr(query), q, ptr(db), n, d, k, int(exclude_self), query_base, int(same), ptr(ws), ws_bytes,
                           ptr(idx), ptr(dist), ptr(stats), ptr(fallback), stream()), "mmu_knn_tc")
    st = stats.tolist()                                   # one small D2H read; the only sync of the call
    n_fb = int(st[0])
    if n_fb:
        # rows the error bound could not certify: exhaustive fp32 search of just those rows
        check(lib().mmu_knn_exact_f32(ptr(query), n_fb, ptr(fallback), ptr(db), n, d, k, int(exclude_self), query_base,
                                      0, 0, ptr(idx), ptr(dist), stream()), "mmu_knn_exact_f32(fallback)")
    last_stats.clear()
    last_stats.update(rows=q, fallback_rows=n_fb, certified_rows=int(st[2]),
                      rescored_per_row=(st[1] / st[2]) if st[2] else 0.0)
    return idx, dist
