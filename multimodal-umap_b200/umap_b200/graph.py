"""Graph construction on the device: exact kNN -> rho/sigma/weights -> fuzzy union.

Host-side orchestration of kernels K1-K6 (include/mmumap.h).  Mirrors what
/root/reference/impl/model.py:63-209 (fuzzy_knn_graph), :271 (union) and :236-252
(embed_query) compute, with the exact kNN search north_star specifies in place of the
reference's randomised NN-descent.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch

from . import dist as D
from . import native, profiler
from .native import check, lib, ptr, stream


@dataclass
class Graph:
    """Coalesced sparse matrix in device CSR+COO form (int32 indices, fp32 values).

    `row`/`col`/`val` are exactly the (row, col)-sorted COO arrays .coalesce() yields in the
    reference (model.py:208,271); `rowptr` (int64, n_rows+1) gives O(1) row ranges instead of
    the reference's full-COO mask scan per batch (model.py:428)."""
    n_rows: int
    n_cols: int
    rowptr: torch.Tensor
    row: torch.Tensor
    col: torch.Tensor
    val: torch.Tensor

    @property
    def nnz(self) -> int:
        return int(self.col.numel())

    def to_sparse_coo(self) -> torch.Tensor:
        idx = torch.stack([self.row.long(), self.col.long()], dim=0)
        # nothing of the engine is attached to the tensor: torch.save pickles a tensor's __dict__, and a checkpoint
        # must stay loadable by the reference / with weights_only=True (impl/model.py keeps a weakref registry instead)
        return torch.sparse_coo_tensor(idx, self.val, (self.n_rows, self.n_cols), is_coalesced=True)

    @staticmethod
    def from_sparse_coo(t: torch.Tensor) -> "Graph":
        native.require_cuda()
        t = t.coalesce() if not t.is_coalesced() else t
        dev = torch.device("cuda")
        idx = t.indices().to(dev)
        row = idx[0].to(torch.int32).contiguous()
        col = idx[1].to(torch.int32).contiguous()
        val = t.values().to(dev, torch.float32).contiguous()
        counts = torch.bincount(idx[0], minlength=t.shape[0])
        rowptr = torch.zeros(t.shape[0] + 1, dtype=torch.int64, device=dev)
        rowptr[1:] = torch.cumsum(counts, 0)
        return Graph(t.shape[0], t.shape[1], rowptr, row, col, val)

    @staticmethod
    def from_fixed_degree(col: torch.Tensor, val: torch.Tensor, n_cols: int) -> "Graph":
        q, k = col.shape
        rowptr = torch.arange(0, (q + 1) * k, k, dtype=torch.int64, device=col.device)
        row = torch.arange(q, dtype=torch.int32, device=col.device).repeat_interleave(k)
        return Graph(q, n_cols, rowptr, row, col.reshape(-1).contiguous(), val.reshape(-1).contiguous())


def _f32c(x: torch.Tensor) -> torch.Tensor:
    native.require_cuda()
    return x.detach().to("cuda", torch.float32).contiguous()


def knn_exact_simt(query: torch.Tensor, db: torch.Tensor, k: int, exclude_self: bool,
                   query_base: int = 0, db_base: int = 0, out=None, rows: torch.Tensor | None = None):
    """K1 fallback path: exhaustive fp32 kNN on CUDA cores (mmu_knn_exact_f32)."""
    query, db = _f32c(query), _f32c(db)
    q, d = query.shape
    merge = out is not None
    if out is None:
        idx = torch.empty((q, k), dtype=torch.int32, device=query.device)
        dist = torch.empty((q, k), dtype=torch.float32, device=query.device)
    else:
        idx, dist = out
    n_query = q if rows is None else int(rows.numel())
    check(lib().mmu_knn_exact_f32(ptr(query), n_query, ptr(rows), ptr(db), db.shape[0], d, k, int(exclude_self),
                                  query_base, db_base, int(merge and rows is None), ptr(idx), ptr(dist), stream()),
          "mmu_knn_exact_f32")
    return idx, dist


def available_knn_methods():
    """kNN engines compiled into the library ("tc" needs the mmu_knn_tc_* entry points)."""
    return ("simt", "tc") if hasattr(lib(), "mmu_knn_tc") else ("simt",)


def knn_graph(query: torch.Tensor, db: torch.Tensor, k: int, exclude_self: bool, method: str | None = None):
    """Exact kNN of `query` rows in `db` (ref: model.py:81-195 role; semantics of
    oracle/knn_oracle.c).  method: "tc" (tcgen05 candidates + fp32 rescoring + certification),
    "simt" (exhaustive fp32 on CUDA cores) or None = env MMUMAP_KNN, default "tc"."""
    method = method or os.environ.get("MMUMAP_KNN", available_knn_methods()[-1])
    if k > native.MAX_K:
        raise ValueError(f"k_neighbors={k} exceeds the supported maximum {native.MAX_K}")
    if db.shape[0] - (1 if exclude_self else 0) < k:
        raise ValueError(f"need more than k={k} candidate points per row, got {db.shape[0]}")
    use_tc = method == "tc" and k <= native.KNN_TC_MAX_K          # the per-row candidate lists hold 64 entries
    if method not in ("tc", "simt"):
        raise ValueError(f"unknown kNN method {method!r}")

    def search(q, d, kk, excl, q_base):
        if use_tc:
            from .knn_tc import knn_tc
            return knn_tc(q, d, kk, excl, query_base=q_base)
        return knn_exact_simt(q, d, kk, excl, query_base=q_base)

    kernel = "knn_tc" if use_tc else "knn_exact_f32_kernel"
    if D.world() > 1 and query is db:
        # multi-GPU fit: query row-blocks per rank, all-gather of the per-row results
        lo, hi = D.row_block(db.shape[0], D.rank(), D.world())
        with profiler.stage("knn", flops=2.0 * (hi - lo) * db.shape[0] * db.shape[1], kernel=kernel):
            if use_tc and exclude_self:
                # large clustered low-dimensional input: the cluster-pruned search, its query blocks sharded over the ranks
                from . import knn_tc as KT
                if KT.prune_applicable(db.shape[0], db.shape[1], k):
                    res = KT.knn_tc(db, db, k, True, prune=True)
                    if res is not None:
                        return res
            if os.environ.get("MMUMAP_KNN_DIST", "rows") == "ring":      # default "rows": impl/model.py DEFAULT_KNN_DIST
                # row-sharded DATABASE: each rank only touches its own rows of `db`; the shards rotate round
                # the ranks peer-to-peer (NCCL send/recv over NVLink) under a running per-row top-k merge
                return knn_ring(_f32c(db)[lo:hi].contiguous(), db.shape[0], k, exclude_self, search)
            return D.knn_sharded_rows(_f32c(db), k, exclude_self, search)
    with profiler.stage("knn", flops=2.0 * query.shape[0] * db.shape[0] * db.shape[1], kernel=kernel):
        return search(query, db, k, exclude_self, 0)


def knn_merge(idx_a: torch.Tensor, dist_a: torch.Tensor, idx_b: torch.Tensor, dist_b: torch.Tensor):
    """K3: merge two sorted per-row lists into one sorted top-k (mmu_knn_merge)."""
    oi, od = torch.empty_like(idx_a), torch.empty_like(dist_a)
    check(lib().mmu_knn_merge(ptr(idx_a.contiguous()), ptr(dist_a.contiguous()), ptr(idx_b.contiguous()),
                              ptr(dist_b.contiguous()), idx_a.shape[0], idx_a.shape[1], ptr(oi), ptr(od), stream()),
          "mmu_knn_merge")
    return oi, od


def knn_ring(x_local: torch.Tensor, n_total: int, k: int, exclude_self: bool, search):
    """Row-sharded database kNN (dist.ring_knn) followed by the all-gather of the row blocks."""
    import torch.distributed as tdist
    li, ld = D.ring_knn(x_local, n_total, k, exclude_self, search, knn_merge)
    w = D.world()
    per = D.block_size(n_total, w)
    ip = torch.full((per, k), -1, dtype=torch.int32, device=x_local.device)
    dp = torch.full((per, k), float("inf"), dtype=torch.float32, device=x_local.device)
    if li is not None:
        ip[: li.shape[0]] = li
        dp[: ld.shape[0]] = ld
    ia = torch.empty((w * per, k), dtype=torch.int32, device=x_local.device)
    da = torch.empty((w * per, k), dtype=torch.float32, device=x_local.device)
    tdist.all_gather_into_tensor(ia, ip)
    tdist.all_gather_into_tensor(da, dp)
    return ia[:n_total].contiguous(), da[:n_total].contiguous()


def smooth_knn(idx: torch.Tensor, dist: torch.Tensor, solver: str = "bisect", n_iter: int | None = None):
    """K4 (ref: model.py:33-61,197-209).  Returns (col_sorted, w_sorted, sigma, rho)."""
    q, k = idx.shape
    code = {"bisect": native.SIGMA_BISECT, "newton": native.SIGMA_NEWTON}[solver]
    if n_iter is None:
        n_iter = 64 if solver == "bisect" else 20
    col = torch.empty_like(idx)
    w = torch.empty_like(dist)
    sigma = torch.empty(q, dtype=torch.float32, device=idx.device)
    rho = torch.empty(q, dtype=torch.float32, device=idx.device)
    check(lib().mmu_smooth_knn(ptr(idx), ptr(dist), q, k, code, n_iter, ptr(sigma), ptr(rho), ptr(col), ptr(w),
                               stream()), "mmu_smooth_knn")
    return col, w, sigma, rho


def invert_weights(idx: torch.Tensor, dist: torch.Tensor, a: float, b: float):
    """ref: model.py:206."""
    q, k = idx.shape
    col = torch.empty_like(idx)
    w = torch.empty_like(dist)
    check(lib().mmu_invert_weights(ptr(idx), ptr(dist), q, k, float(a), float(b), ptr(col), ptr(w), stream()),
          "mmu_invert_weights")
    return col, w


# multi-GPU: graphs at least this large build the union one row block per rank.  Measured at 4 GPUs on 1M x 15 entries the
# host glue of the sharded form (filter, padded all-gathers, concatenation) costs 34 ms against 4 ms for the replicated
# single-GPU kernel chain, so only C4-class graphs (10M x 30: 86 ms replicated) are sharded.
UNION_SHARD_MIN_EDGES = 100_000_000


def fuzzy_union_rows(col: torch.Tensor, w: torch.Tensor, lo: int, hi: int):
    """Rows [lo, hi) of the union as (rowptr_local int64 [hi-lo+1], row, col, val): the block's own rows of G plus the
    in-edges of the block filtered out of the replicated graph (source-major order), mmu_fuzzy_union_rows."""
    n, k = col.shape
    dev = col.device
    flat = col.reshape(-1)
    e = ((flat >= lo) & (flat < hi)).nonzero(as_tuple=True)[0]                  # ascending entry number = source-major
    in_key = (flat.index_select(0, e) - lo).to(torch.int32).contiguous()
    in_src = (e // k).to(torch.int32).contiguous()
    in_w = w.reshape(-1).index_select(0, e).contiguous()
    n_in = int(e.numel())
    nb = hi - lo
    ws_bytes = lib().mmu_union_rows_workspace_bytes(nb, n_in)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    cap = nb * k + n_in
    rowptr = torch.empty(nb + 1, dtype=torch.int64, device=dev)
    orow = torch.empty(cap, dtype=torch.int32, device=dev)
    ocol = torch.empty(cap, dtype=torch.int32, device=dev)
    oval = torch.empty(cap, dtype=torch.float32, device=dev)
    cb, wb = col[lo:hi].contiguous(), w[lo:hi].contiguous()
    check(lib().mmu_fuzzy_union_rows(ptr(cb), ptr(wb), nb, k, ptr(in_key), ptr(in_src), ptr(in_w), n_in, lo, ptr(ws), ws_bytes,
                                     ptr(rowptr), ptr(orow), ptr(ocol), ptr(oval), stream()), "mmu_fuzzy_union_rows")
    nnz = int(rowptr[-1].item())
    return rowptr, orow[:nnz], ocol[:nnz], oval[:nnz]


def fuzzy_union_sharded(col: torch.Tensor, w: torch.Tensor) -> Graph:
    """Multi-GPU union: rank r builds the rows row_block(r) of S -- the exchange of edges keyed by destination
    row block is the filter of the replicated kNN result -- and the CSR blocks are all-gathered (padded to the largest)."""
    import torch.distributed as tdist
    n, k = col.shape
    dev = col.device
    world, rank = D.world(), D.rank()
    lo, hi = D.row_block(n, rank, world)
    if hi > lo:
        rp, r_, c_, v_ = fuzzy_union_rows(col, w, lo, hi)
    else:
        rp = torch.zeros(1, dtype=torch.int64, device=dev)
        r_ = torch.zeros(0, dtype=torch.int32, device=dev)
        c_, v_ = r_.clone(), torch.zeros(0, dtype=torch.float32, device=dev)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([r_.numel()], dtype=torch.int64, device=dev)
    tdist.all_gather_into_tensor(counts, mine)
    counts_h = counts.tolist()
    mx = max(max(counts_h), 1)

    def gather(t):
        pad = torch.zeros(mx, dtype=t.dtype, device=dev)
        pad[: t.numel()] = t
        full = torch.empty(world * mx, dtype=t.dtype, device=dev)
        tdist.all_gather_into_tensor(full, pad)
        return torch.cat([full[i * mx: i * mx + counts_h[i]] for i in range(world)])

    row, colo, val = gather(r_), gather(c_), gather(v_)
    per = D.block_size(n, world)
    cnt_pad = torch.zeros(per, dtype=torch.int64, device=dev)
    cnt_pad[: hi - lo] = rp[1:] - rp[:-1]
    cnt_all = torch.empty(world * per, dtype=torch.int64, device=dev)
    tdist.all_gather_into_tensor(cnt_all, cnt_pad)
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    rowptr[1:] = torch.cumsum(cnt_all[:n], 0)
    return Graph(n, n, rowptr, row, colo, val)


def fuzzy_union(col: torch.Tensor, w: torch.Tensor) -> Graph:
    """K5 (ref: model.py:271): S = G + G^T - G*G^T for the fixed-degree graph (col, w) [n x k]."""
    n, k = col.shape
    dev = col.device
    if D.world() > 1 and n * k >= UNION_SHARD_MIN_EDGES and os.environ.get("MMUMAP_UNION_SHARD", "1") == "1":
        return fuzzy_union_sharded(col, w)
    ws_bytes = lib().mmu_union_workspace_bytes(n, k)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    cap = 2 * n * k
    rowptr = torch.empty(n + 1, dtype=torch.int64, device=dev)
    orow = torch.empty(cap, dtype=torch.int32, device=dev)
    ocol = torch.empty(cap, dtype=torch.int32, device=dev)
    oval = torch.empty(cap, dtype=torch.float32, device=dev)
    check(lib().mmu_fuzzy_union(ptr(col), ptr(w), n, k, ptr(ws), ws_bytes, ptr(rowptr), ptr(orow), ptr(ocol),
                                ptr(oval), stream()), "mmu_fuzzy_union")
    nnz = int(rowptr[-1].item())
    return Graph(n, n, rowptr, orow[:nnz].clone(), ocol[:nnz].clone(), oval[:nnz].clone())


def embed_query(col: torch.Tensor, w: torch.Tensor, ref: torch.Tensor) -> torch.Tensor:
    """K6 (ref: model.py:236-252)."""
    q, k = col.shape
    ref = _f32c(ref)
    out = torch.empty((q, ref.shape[1]), dtype=torch.float32, device=col.device)
    check(lib().mmu_embed_query(ptr(col), ptr(w), q, k, ptr(ref), ref.shape[1], ptr(out), stream()),
          "mmu_embed_query")
    return out


def spmm(g: Graph, x: torch.Tensor, val: torch.Tensor | None = None) -> torch.Tensor:
    x = x.contiguous()
    y = torch.empty((g.n_rows, x.shape[1]), dtype=torch.float32, device=x.device)
    v = g.val if val is None else val
    check(lib().mmu_spmm_csr(ptr(g.rowptr), ptr(g.col), ptr(v), g.n_rows, ptr(x), x.shape[1], ptr(y), stream()),
          "mmu_spmm_csr")
    return y


def spmm_axpby(g: Graph, val: torch.Tensor, x: torch.Tensor, alpha: float, beta: float, z: torch.Tensor | None,
               gamma: float, out: torch.Tensor | None = None) -> torch.Tensor:
    """out = alpha * (A x) + beta * x + gamma * z with A = (g pattern, val); z may be `out`."""
    if out is None:
        out = torch.empty_like(x)
    check(lib().mmu_spmm_csr_axpby(ptr(g.rowptr), ptr(g.col), ptr(val), g.n_rows, ptr(x), x.shape[1], float(alpha),
                                   float(beta), ptr(z), float(gamma), ptr(out), stream()), "mmu_spmm_csr_axpby")
    return out
