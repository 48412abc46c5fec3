"""Single-node multi-GPU partitioning (one process per GPU, torch.distributed).

The reference is single-process; this is the partitioning SURVEY.md 8(e)
derives for its path:

  kNN        independent query rows -> each rank searches a contiguous block of query rows
             (aligned to the 128-row CTA tile) and the (idx, dist) blocks are all-gathered.  With a
             row-sharded database (`ring_knn`) the database shards rotate round the ranks
             peer-to-peer (NCCL send/recv over NVLink) while each rank keeps a running per-row
             top-k that is merged shard by shard (mmu_knn_merge).
  sigma/rho, fuzzy union
             replicated on the all-gathered kNN result (sub-millisecond, deterministic).
  spectral   modality m is solved by rank m % world and broadcast.
  optimiser  edges sharded by row-batch ranges (so every batch's kept count is local), InfoNCE
             anchors sharded by range; embeddings and Adam state replicated; ONE all-reduce of the
             flat gradient buffer per epoch, then the identical Adam step on every rank.  The
             random streams are keyed on global edge / anchor positions, so N ranks draw exactly
             the single-GPU sample stream.

Nothing here touches CUDA directly: the compute steps are passed in as callables, which is what
lets the host logic be tested with the gloo backend on CPU (tests/test_dist_cpu.py).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


_LOCAL_ONLY = 0


def world() -> int:
    if _LOCAL_ONLY:
        return 1
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    if _LOCAL_ONLY:
        return 0
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


class local_only:
    """Inside this context the partitioning helpers see ONE rank: the enclosed work is this rank's own shard of an
    embarrassingly parallel stage (the transform of independent query rows) and runs the single-GPU path, with no
    exchange.  real_world() / real_rank() still report the process group."""

    def __enter__(self):
        global _LOCAL_ONLY
        _LOCAL_ONLY += 1
        return self

    def __exit__(self, *exc):
        global _LOCAL_ONLY
        _LOCAL_ONLY -= 1
        return False


def real_world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def real_rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def all_gather_rows(local: torch.Tensor, n_total: int, per: int) -> torch.Tensor:
    """Concatenate the ranks' row blocks (rank r holds rows [r*per, r*per + local.shape[0])) into the full
    [n_total, ...] tensor on every rank."""
    w = real_world()
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    full = torch.empty((w * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(full, pad)
    return full[:n_total].contiguous()


def row_block(n: int, r: int, w: int, align: int = 128) -> tuple[int, int]:
    """Contiguous block of rows of rank r: equal blocks of ceil(n/w) rounded up to `align`."""
    per = -(-n // w)
    per = -(-per // align) * align
    lo = min(n, r * per)
    return lo, min(n, lo + per)


def block_size(n: int, w: int, align: int = 128) -> int:
    per = -(-n // w)
    return -(-per // align) * align


def batch_range(n_batches: int, r: int, w: int) -> tuple[int, int]:
    """Row-batches (model.py:423) owned by rank r."""
    return (n_batches * r) // w, (n_batches * (r + 1)) // w


def balanced_batch_range(cum_weight: list, r: int, w: int) -> tuple[int, int]:
    """Row-batches owned by rank r when the shards are balanced by WEIGHT instead of by count: cum_weight[b] is
    the sum of the edge weights of batches 0..b (the expected number of kept edges, model.py:432).  Rank r gets
    the batches whose cumulative weight ends in (total*r/w, total*(r+1)/w]: contiguous, disjoint, covering, and
    identical on every rank (a pure function of the replicated graph)."""
    import bisect
    nb = len(cum_weight)
    total = cum_weight[-1] if nb else 0.0
    if total <= 0.0:
        return batch_range(nb, r, w)
    cuts = [0] * (w + 1)
    cuts[w] = nb
    for i in range(1, w):
        cuts[i] = min(nb, max(cuts[i - 1], bisect.bisect_left(cum_weight, total * i / w) + 1))
    return cuts[r], cuts[r + 1]


def item_range(n: int, r: int, w: int) -> tuple[int, int]:
    return (n * r) // w, (n * (r + 1)) // w


def knn_sharded_rows(x: torch.Tensor, k: int, exclude_self: bool, knn_fn):
    """Exact kNN of every row of `x` (replicated on all ranks) in `x`: rank r searches rows
    row_block(r) with knn_fn(query_rows, db, k, exclude_self, query_base) and the blocks are
    all-gathered.  Returns the full (idx [n,k] int32, dist [n,k] float32) on every rank."""
    w, r = world(), rank()
    n = x.shape[0]
    if w == 1:
        return knn_fn(x, x, k, exclude_self, 0)
    lo, hi = row_block(n, r, w)
    per = block_size(n, w)
    idx_pad = torch.full((per, k), -1, dtype=torch.int32, device=x.device)
    dist_pad = torch.full((per, k), float("inf"), dtype=torch.float32, device=x.device)
    from . import profiler
    if hi > lo:
        with profiler.stage("knn_local"):
            idx_l, dist_l = knn_fn(x[lo:hi], x, k, exclude_self, lo)
        idx_pad[: hi - lo] = idx_l
        dist_pad[: hi - lo] = dist_l
    idx_all = torch.empty((w * per, k), dtype=torch.int32, device=x.device)
    dist_all = torch.empty((w * per, k), dtype=torch.float32, device=x.device)
    with profiler.stage("knn_allgather"):
        dist.all_gather_into_tensor(idx_all, idx_pad)
        dist.all_gather_into_tensor(dist_all, dist_pad)
    return idx_all[:n].contiguous(), dist_all[:n].contiguous()


def ring_knn(x_local: torch.Tensor, n_total: int, k: int, exclude_self: bool, knn_fn, merge_fn):
    """Exact kNN with the DATABASE sharded by rows: rank r owns rows row_block(r) = x_local.
    Shards rotate round the ring (rank r sends the shard it holds to r+1 and receives from r-1)
    while the search against the shard in hand runs; knn_fn(query, db, k, exclude_self, query_base)
    returns db-local indices, merge_fn(idx_a, dist_a, idx_b, dist_b) merges two sorted lists.
    Returns (idx, dist) of the local rows with GLOBAL database indices."""
    w, r = world(), rank()
    per = block_size(n_total, w)
    d = x_local.shape[1]
    q = x_local.shape[0]
    dev = x_local.device
    held = torch.zeros((per, d), dtype=x_local.dtype, device=dev)
    held[:q] = x_local
    held_rank = r
    best_i = best_d = None
    for step in range(w):
        lo, hi = row_block(n_total, held_rank, w)
        nxt = None
        reqs = []
        if step + 1 < w:
            nxt = torch.empty_like(held)
            ops = [dist.P2POp(dist.isend, held, (r + 1) % w), dist.P2POp(dist.irecv, nxt, (r - 1) % w)]
            reqs = dist.batch_isend_irecv(ops)
        if hi > lo and q > 0:
            same = held_rank == r
            i_s, d_s = knn_fn(x_local, held[: hi - lo], k, exclude_self and same, 0)
            i_s = torch.where(i_s >= 0, i_s + lo, i_s)
            if best_i is None:
                best_i, best_d = i_s, d_s
            else:
                best_i, best_d = merge_fn(best_i, best_d, i_s, d_s)
        for req in reqs:
            req.wait()
        if nxt is not None:
            held = nxt
            held_rank = (held_rank - 1) % w
    return best_i, best_d


def all_reduce_sum(t: torch.Tensor) -> None:
    if world() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)


def broadcast(t: torch.Tensor, src: int) -> None:
    if world() > 1:
        dist.broadcast(t, src=src)


def same_on_all_ranks(value: int) -> int:
    """Rank 0's value everywhere (seeds must agree: the embeddings are replicated)."""
    if world() == 1:
        return value
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([value], dtype=torch.int64, device=dev)
    dist.broadcast(t, src=0)
    return int(t.item())
