"""Tensor-core exact kNN (K1/K2, include/mmumap.h: mmu_knn_tc).

Role of /root/reference/impl/model.py:81-195 (candidate search + per-row top-k).  The tcgen05
kernel generates candidates, every row is certified and rescored with the canonical fp32
distance.  Rows the first pass cannot certify are retried with a deeper candidate pool (the
database searched in 8 splits of 64 candidates each) and error-compensated split-fp16 operands
(hi + lo: error bound 2^-20 |x||y| instead of 2^-9 |x||y|, three times the MMA work); what is still
uncertified (exact ties beyond the pool, degenerate data) is finished by the exhaustive fp32
kernel, so the result equals mmu_knn_exact_f32 bit for bit.
"""
from __future__ import annotations

import torch

from . import dist as D
from . import native
from .native import check, lib, ptr, stream

last_stats: dict = {}
DEEP_SPLITS = 8
PRUNE_MIN_ROWS = 262144        # below this the full contraction takes milliseconds
PRUNE_MAX_DIM = 256
PRUNE_CONTRAST = 0.4           # k-th neighbour distance / distance between random rows (knn_pruned.contrast)


def prune_applicable(n_rows: int, dim: int, k: int) -> bool:
    import os
    return (n_rows >= PRUNE_MIN_ROWS and dim <= PRUNE_MAX_DIM and k + 1 <= 64
            and os.environ.get("MMUMAP_KNN_PRUNE", "1") == "1")


def _call(query, db, k, exclude_self, query_base, gid, same, min_splits, precision=0):
    q, d = query.shape
    n = db.shape[0]
    dev = db.device
    ws_bytes = lib().mmu_knn_tc_workspace_bytes(q, n, d, int(same), min_splits, precision)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    idx = torch.empty((q, k), dtype=torch.int32, device=dev)
    dist = torch.empty((q, k), dtype=torch.float32, device=dev)
    stats = torch.empty(4, dtype=torch.int32, device=dev)
    fallback = torch.empty(max(q, 1), dtype=torch.int32, device=dev)
    check(lib().mmu_knn_tc(ptr(query), q, ptr(db), n, d, k, int(exclude_self), query_base, ptr(gid), int(same), min_splits,
                           precision, ptr(ws), ws_bytes, ptr(idx), ptr(dist), ptr(stats), ptr(fallback), stream()), "mmu_knn_tc")
    st = stats.tolist()                                   # one small D2H read; the only sync of the call
    return idx, dist, st, fallback


def knn_tc(query: torch.Tensor, db: torch.Tensor, k: int, exclude_self: bool, query_base: int = 0, prune: bool | None = None):
    """Returns (idx int32 [Q,k] sorted by (dist, idx), dist float32 [Q,k]).  prune: None = the cluster-pruned search
    (knn_pruned.py) where it applies and pays, False = never, True = pruned or nothing (returns None when it does not
    apply: the multi-GPU caller then runs its row-sharded full search instead of a replicated one)."""
    native.require_cuda()
    if k > native.KNN_TC_MAX_K:
        raise ValueError(f"tensor-core kNN supports k <= {native.KNN_TC_MAX_K}, got {k}")
    same = query is db or (query.data_ptr() == db.data_ptr() and query.shape == db.shape)
    db = db.detach().to("cuda", torch.float32).contiguous()
    query = db if same else query.detach().to("cuda", torch.float32).contiguous()
    same = same and query_base == 0
    q = query.shape[0]
    n = db.shape[0]
    # candidate pool depth: one 64-entry list certifies k <= 16 comfortably; deeper pools for larger k
    min_splits = 0 if k <= 16 else 2
    precision = 0
    pruned = None
    if prune is not False and same and exclude_self and prune_applicable(q, query.shape[1], k):
        # large, low-dimensional input: if its neighbourhoods are small against the typical distance between rows
        # (clustered data), search cluster by cluster with exact ball bounds instead of the full contraction
        from . import knn_pruned
        ratio = knn_pruned.contrast(db, k)
        if ratio < PRUNE_CONTRAST:
            pruned = knn_pruned.knn_pruned(db, k)
    if prune is True and pruned is None:
        return None
    if pruned is not None:
        from . import knn_pruned
        idx, dist, fb_rows = pruned
        n_fb = n_first = int(fb_rows.numel())
        fallback = fb_rows.to(torch.int32)
        certified = q - n_fb
        rescored = int(knn_pruned.last_stats.get("rescored_per_row", 0.0) * certified)
        min_splits, precision = 0, 1
        last_stats_extra = {"pruning": dict(knn_pruned.last_stats, contrast=ratio)}
    elif q >= 65536 and n >= 65536 and (k > 16 or query.shape[1] <= 256):
        # probe a strided sample of rows: when the data's neighbourhood gaps are small against the fp16
        # error bound (low dimension, large N, large norms), go straight to the deep pool for all rows
        step = q // 2048
        rows = torch.arange(0, q, step, device=db.device, dtype=torch.int64)[:2048]
        gid = (rows + query_base).to(torch.int32)
        _, _, pst, _ = _call(query.index_select(0, rows), db, k, exclude_self, 0, gid, False, min_splits)
        if pst[0] > 0.1 * rows.numel():
            precision = 1          # split operands shrink the error bound ~2000x: the shallow pool suffices again
    if pruned is None:
        idx, dist, st, fallback = _call(query, db, k, exclude_self, query_base, None, same, min_splits, precision)
        n_fb = n_first = int(st[0])
        rescored, certified = st[1], st[2]
        last_stats_extra = {}
    shard_fb = None
    if pruned is not None and n_fb and D.world() > 1:
        # multi-GPU: the rows the pruned pass left open are known to every rank; each finishes every W-th of them
        shard_fb = fallback[:n_fb].long()
        fallback = shard_fb[D.rank()::D.world()].to(torch.int32).contiguous()
        n_fb = int(fallback.numel())
    if n_fb and not (precision == 1 and min_splits >= DEEP_SPLITS):
        # second level: only the uncertified rows, database in 8 splits (512 candidates per row), split operands
        rows = fallback[:n_fb].long()
        gid = (rows + query_base).to(torch.int32)
        idx2, dist2, st2, fb2 = _call(query.index_select(0, rows), db, k, exclude_self, 0, gid, False, DEEP_SPLITS, 1)
        idx.index_copy_(0, rows, idx2)
        dist.index_copy_(0, rows, dist2)
        n_fb = int(st2[0])
        rescored, certified = rescored + st2[1], certified + st2[2]
        fallback = rows.index_select(0, fb2[:n_fb].long()).to(torch.int32) if n_fb else fallback
    if n_fb:
        # rows no candidate pool can certify (exact ties beyond the pool, degenerate data): exhaustive fp32
        fallback = fallback[:n_fb].contiguous()
        check(lib().mmu_knn_exact_f32(ptr(query), n_fb, ptr(fallback), ptr(db), n, query.shape[1], k, int(exclude_self),
                                      query_base, 0, 0, ptr(idx), ptr(dist), stream()), "mmu_knn_exact_f32(fallback)")
    if shard_fb is not None:
        # exchange the finished rows: pad every rank's share to the same length and all-gather
        import torch.distributed as tdist
        w = D.world()
        per = -(-int(shard_fb.numel()) // w)
        mine = shard_fb[D.rank()::w]
        pad_rows = torch.full((per,), -1, dtype=torch.int64, device=idx.device)
        pad_rows[: mine.numel()] = mine
        pad_i = torch.zeros((per, k), dtype=torch.int32, device=idx.device)
        pad_d = torch.zeros((per, k), dtype=torch.float32, device=idx.device)
        pad_i[: mine.numel()] = idx.index_select(0, mine)
        pad_d[: mine.numel()] = dist.index_select(0, mine)
        all_rows = torch.empty(w * per, dtype=torch.int64, device=idx.device)
        all_i = torch.empty((w * per, k), dtype=torch.int32, device=idx.device)
        all_d = torch.empty((w * per, k), dtype=torch.float32, device=idx.device)
        tdist.all_gather_into_tensor(all_rows, pad_rows)
        tdist.all_gather_into_tensor(all_i, pad_i)
        tdist.all_gather_into_tensor(all_d, pad_d)
        ok = all_rows >= 0
        idx.index_copy_(0, all_rows[ok], all_i[ok])
        dist.index_copy_(0, all_rows[ok], all_d[ok])
        t = torch.tensor([n_fb], dtype=torch.int64, device=idx.device)
        D.all_reduce_sum(t)
        n_fb = int(t.item())
    last_stats.clear()
    last_stats.update(rows=q, first_pass_uncertified=n_first, fallback_rows=n_fb, certified_rows=int(certified),
                      rescored_per_row=(rescored / certified) if certified else 0.0, min_splits=min_splits,
                      precision=precision, pruned=pruned is not None, **last_stats_extra)
    return idx, dist
