"""Exact kNN with cluster pruning for large, low-dimensional, clustered inputs (BASELINE.json configs[3]: 10M x 128).

Role of /root/reference/impl/model.py:81-195 (candidate search + per-row top-k), same result contract as
knn_tc.knn_tc: indices and fp32 distance bit patterns of the exhaustive search.  On data of this kind the dense
contraction is the wrong algorithm at full size -- 10^14 distances of which all but a sliver are between points of
different clusters -- so the database is re-ordered by cluster and every 128-row query block searches only the 256-row
tiles that can hold one of its neighbours:

  1. centroids: farthest-point sampling on a strided subsample (every cluster of the data gets one);
  2. assignment: each row's nearest centroid and its fp32 distance to it, with the engine's own kNN kernel (k = 1);
  3. rows sorted by cluster (`perm`), cluster radius R_b = largest member distance;
  4. pass 1 (mmu_knn_tc_ex, per-block tile RANGE): every query block searches the tiles of its own cluster(s);
  5. bound: U = the block's largest (k+1)-th smallest approximate score + error bound >= every row's true k-th
     neighbour distance (squared); a cluster b must still be searched iff
         |c_a - c_b| - rho_block - R_b  <  sqrt(U)          (triangle inequality; rho_block = largest |x - c_a| in the block)
  6. pass 2 (per-block tile LIST, candidate lists resumed): the tiles of those clusters not yet visited;
  7. the usual certification + canonical fp32 rescoring, reporting ORIGINAL indices (db_gid = perm), so ties and self
     exclusion are decided exactly as in the unpruned search.

A row is exact because (a) inside the visited tiles the certification of knn_tc.cu holds unchanged and (b) every point of
an unvisited tile is farther than the row's k-th neighbour by the bound of step 5 (taken with a 1e-3 relative safety
margin against the fp32 rounding of the centroid distances and radii).  Rows the certification rejects go through the
same second level / exhaustive fallback as in knn_tc.knn_tc.
"""
from __future__ import annotations

import ctypes
import math
import os

import torch

from . import dist as D
from . import native
from .native import check, lib, ptr, stream

last_stats: dict = {}
SAFETY = 1e-3
ONE_SPLIT = -2                 # min_splits value that PINS the split count: two splits take a block's tiles in turn, so a row's
                               # neighbours are spread over 2 x 4 lists of 16 (by tile parity and column slice); with one split
                               # 0.6 % of the rows had more than 16 of their 30 neighbours in one list and could not certify


def _spread_rows(n: int, count: int, device) -> torch.Tensor:
    """`count` distinct row numbers spread over [0, n) by multiplicative hashing: deterministic (every rank picks the same
    rows, no RNG state consumed) and -- unlike a fixed stride -- not in resonance with periodic row orders (synthetic data
    laid out cluster by cluster modulo the cluster count: a stride of 15 over 1000 clusters only ever sees 200 of them)."""
    count = min(count, n)
    i = torch.arange(count, device=device, dtype=torch.int64)
    rows = (i * 2654435761 + 12345) % n
    return torch.unique(rows) if count < n else torch.arange(n, device=device, dtype=torch.int64)


def contrast(x: torch.Tensor, k: int, sample: int = 1024) -> float:
    """median k-th neighbour distance of a strided row sample / median distance between random rows: small when the
    data is clustered at the scale of its neighbourhoods, i.e. when ball bounds can prune."""
    from .knn_tc import _call
    n = x.shape[0]
    rows = _spread_rows(n, sample, x.device)
    _, dist, st, fb = _call(x.index_select(0, rows), x, k, True, 0, rows.to(torch.int32), False, 0, 1)
    ok = torch.ones(rows.numel(), dtype=torch.bool, device=x.device)
    ok[fb[: int(st[0])].long()] = False                     # uncertified rows are not written by the call
    ok &= torch.isfinite(dist[:, k - 1]) & (dist[:, k - 1] > 0)
    if int(ok.sum()) < rows.numel() // 2:
        return 1.0
    dk = dist[ok, k - 1].median()
    other = x.index_select(0, (rows * 7919 + 13) % n)
    dr = (x.index_select(0, rows) - other).norm(dim=1).median()
    return float((dk / dr.clamp(min=1e-30)).item())


def farthest_point_centroids(x: torch.Tensor, n_centroids: int, sub_rows: int = 65536) -> torch.Tensor:
    """Greedy k-centre on a hashed subsample (mmu_fps_centroids: one persistent kernel, a grid barrier per round):
    deterministic -- no RNG, every rank gets the same centroids."""
    n = x.shape[0]
    sub = x if n <= sub_rows else x.index_select(0, _spread_rows(n, sub_rows, x.device))
    sub = sub.contiguous()
    n_centroids = min(n_centroids, sub.shape[0])
    L = lib()
    ws = torch.empty(L.mmu_fps_workspace_bytes(sub.shape[0]), dtype=torch.uint8, device=x.device)
    cent = torch.empty((n_centroids, x.shape[1]), dtype=torch.float32, device=x.device)
    rows = torch.empty(n_centroids, dtype=torch.int32, device=x.device)
    check(L.mmu_fps_centroids(ptr(sub), sub.shape[0], x.shape[1], n_centroids, ptr(ws), ws.numel(), ptr(cent), ptr(rows), stream()),
          "mmu_fps_centroids")
    return cent


def farthest_point_centroids_torch(x: torch.Tensor, n_centroids: int, sub_rows: int = 65536) -> torch.Tensor:
    """the same selection with torch ops (test reference for the kernel)"""
    n = x.shape[0]
    sub = x if n <= sub_rows else x.index_select(0, _spread_rows(n, sub_rows, x.device))
    d2 = torch.full((sub.shape[0],), float("inf"), device=x.device)
    cent = torch.empty((n_centroids, x.shape[1]), dtype=torch.float32, device=x.device)
    i = torch.zeros((), dtype=torch.int64, device=x.device)
    for c in range(n_centroids):
        row = sub.index_select(0, i.reshape(1))
        cent[c] = row[0]
        d2 = torch.minimum(d2, (sub - row).square().sum(dim=1))
        i = d2.argmax()
    return cent


def _views(ws: torch.Tensor, words, n_rows_pad: int):
    off_prm, off_xnorm, _, _, off_cscore, _, n_qb, n_splits, _, kp, bm, _ = words
    prm = ws[off_prm:off_prm + 16].view(torch.float32)
    xnorm = ws[off_xnorm:off_xnorm + 4 * n_qb * bm].view(torch.float32)
    cscore = ws[off_cscore:off_cscore + 4 * n_qb * n_splits * kp * bm].view(torch.float32).view(n_qb, n_splits * kp, bm)
    return prm, xnorm, cscore


def knn_pruned(x: torch.Tensor, k: int, n_centroids: int | None = None, max_fraction: float = 0.5):
    """Exact kNN of every row of `x` among the rows of `x` (self excluded).  Returns (idx int32 [N,k], dist float32 [N,k],
    fallback_rows int64 [F]) with the rows of `fallback_rows` NOT filled in (the caller runs them through the deeper
    levels of knn_tc), or None when the tile lists show that pruning does not pay (the caller runs the full search)."""
    native.require_cuda()
    L = lib()
    st = stream()
    dev = x.device
    n, dim = x.shape
    world, rank = D.world(), D.rank()
    debug = os.environ.get("MMUMAP_KNN_DEBUG") == "1"
    marks = []

    def mark(label):
        if debug:
            import time
            torch.cuda.synchronize()
            marks.append((label, time.perf_counter()))

    mark("start")
    if n_centroids is None:
        n_centroids = int(min(8192, max(256, 2 ** round(math.log2(1.3 * math.sqrt(n))))))
    # ---- 1-3: centroids, assignment, cluster order
    cent = farthest_point_centroids(x, n_centroids)
    mark("centroids (farthest-point sampling)")
    from .knn_tc import knn_tc
    if world > 1:
        # every rank assigns a block of rows; the blocks are all-gathered
        lo, hi = D.row_block(n, rank, world)
        a_i, a_d = knn_tc(x[lo:hi].contiguous(), cent, 1, False) if hi > lo else (
            torch.zeros((0, 1), dtype=torch.int32, device=dev), torch.zeros((0, 1), dtype=torch.float32, device=dev))
        per = D.block_size(n, world)
        a_idx = D.all_gather_rows(a_i, n, per)
        a_dist = D.all_gather_rows(a_d, n, per)
    else:
        a_idx, a_dist = knn_tc(x, cent, 1, False)           # nearest centroid and the canonical fp32 distance to it
    a_idx = a_idx[:, 0].long()
    a_dist = a_dist[:, 0]
    mark("assignment (kNN k=1 against the centroids)")
    perm = torch.argsort(a_idx, stable=True)
    # Inside every full 256-row tile the sorted rows are dealt round-robin to the tile's four 64-column slices
    # (row j of the tile goes to slot (j mod 4) * 64 + j div 4).  The candidates kernel keeps one 16-entry list per row
    # and COLUMN SLICE (knn_tc.cu, TC_EPI_GROUPS): in cluster order a row's neighbours sit in a few adjacent columns,
    # one slice would have to hold most of the top k and could not certify; dealt out, every slice sees a quarter.
    full = (n // 256) * 256
    if full:
        j = torch.arange(256, device=dev)
        slot_src = torch.empty(256, dtype=torch.int64, device=dev)
        slot_src[(j % 4) * 64 + j // 4] = j                       # slot -> source row of the tile
        perm[:full] = perm[:full].view(-1, 256).index_select(1, slot_src).reshape(-1)
    assign = a_idx.index_select(0, perm)
    counts = torch.bincount(a_idx, minlength=n_centroids)
    ends = torch.cumsum(counts, 0)
    starts = ends - counts
    radius = torch.zeros(n_centroids, device=dev).scatter_reduce(0, a_idx, a_dist, reduce="amax", include_self=True)
    xs = x.index_select(0, perm)
    perm32 = perm.to(torch.int32)
    # ---- query rows of this rank: a block of whole 128-row query blocks of the SORTED order
    q_lo, q_hi = D.row_block(n, rank, world) if world > 1 else (0, n)
    xq = xs if world == 1 else xs[q_lo:q_hi].contiguous()
    gq = perm32 if world == 1 else perm32[q_lo:q_hi].contiguous()
    nq = xq.shape[0]
    same = world == 1
    precision = 1                      # split-fp16 operands: after pruning the contraction is cheap, certification is not
    ws_bytes = L.mmu_knn_tc_workspace_bytes(nq, n, dim, int(same), ONE_SPLIT, precision)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    words = (ctypes.c_int64 * 12)()
    consts = (ctypes.c_float * 4)()
    check(L.mmu_knn_tc_layout(nq, n, dim, int(same), ONE_SPLIT, precision, words, consts), "mmu_knn_tc_layout")
    words = list(words)
    n_qb, n_splits, n_tiles, kp, bm, bn = words[6], words[7], words[8], words[9], words[10], words[11]
    if n_splits != -ONE_SPLIT or k + 1 > kp:
        return None
    idx = torch.empty((nq, k), dtype=torch.int32, device=dev)
    dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
    stats = torch.zeros(4, dtype=torch.int32, device=dev)
    fallback = torch.empty(max(nq, 1), dtype=torch.int32, device=dev)

    def stage(mask, tb=None, te=None, tp=None, tl=None, resume=0):
        check(L.mmu_knn_tc_ex(ptr(xq), nq, ptr(xs), n, dim, k, 1, 0, ptr(gq), int(same), ONE_SPLIT, precision, ptr(ws), ws_bytes,
                              ptr(idx), ptr(dist), ptr(stats), ptr(fallback), mask, ptr(tb), ptr(te), ptr(tp), ptr(tl), resume,
                              ptr(perm32), st), "mmu_knn_tc_ex")

    mark("cluster order, permuted copy")
    stage(1)
    mark("prep (fp16 split operands)")
    # ---- 4: pass 1 over the home clusters' tiles (one contiguous range per query block)
    blk_assign = torch.full((n_qb * bm,), -1, dtype=torch.int64, device=dev)
    blk_assign[:nq] = assign[q_lo:q_hi]
    a1 = blk_assign.view(n_qb, bm).amax(dim=1)                                          # largest / smallest cluster id in the block
    blk_assign[nq:] = n_centroids
    a0 = blk_assign.view(n_qb, bm).amin(dim=1)
    t_begin = (starts.index_select(0, a0) // bn).to(torch.int32)
    t_end = ((ends.index_select(0, a1) + bn - 1) // bn).to(torch.int32)
    stage(2, tb=t_begin, te=t_end)
    mark("pass 1 (home clusters)")
    # ---- 5: bound.  U (scaled, squared) >= true k-th neighbour distance of every row of the block
    prm, xnorm, cscore = _views(ws, words, nq)
    scale = prm[0]
    ymax2 = prm[1]                                           # largest |Y|^2 of the scaled database (float bits)
    c_rel, c_norm, c_abs, _ = (float(v) for v in consts)
    kth = torch.kthvalue(cscore, k + 1, dim=1).values                                  # [n_qb, 128]; +inf when the list is short
    x2 = xnorm[: n_qb * bm].view(n_qb, bm)
    eps = c_rel * x2.sqrt() * ymax2.sqrt() + c_norm * ymax2 + c_abs
    u = (kth + eps + x2).clamp(min=0.0)
    valid_row = (torch.arange(n_qb * bm, device=dev).view(n_qb, bm) + q_lo) < q_hi
    u = torch.where(valid_row, u, torch.zeros_like(u))
    # A 128-row query block of the sorted order usually straddles clusters (clusters are ~N/C rows, and neighbouring
    # cluster ids are unrelated places), so the bound is taken per SEGMENT = the rows of one cluster inside one block:
    # centre = the cluster's centroid, rho_seg = largest member distance |x - c_a| among the segment's rows (known exactly
    # from the assignment), r_seg = largest k-th-neighbour bound among them.  A block searches the union of its segments' needs.
    rows_local = torch.arange(nq, device=dev)
    row_u = u.reshape(-1)[:nq]
    row_rk = row_u.sqrt() / scale                                                      # original units, per row
    row_a = assign[q_lo:q_hi]
    row_rho = a_dist.index_select(0, perm[q_lo:q_hi])
    seg_key = (rows_local // bm) * n_centroids + row_a
    _, seg_of_row = torch.unique_consecutive(seg_key, return_inverse=True)
    n_seg = int(seg_of_row[-1].item()) + 1 if nq else 0
    seg_rk = torch.zeros(n_seg, device=dev).scatter_reduce(0, seg_of_row, row_rk, reduce="amax", include_self=True)
    seg_rho = torch.zeros(n_seg, device=dev).scatter_reduce(0, seg_of_row, row_rho, reduce="amax", include_self=True)
    seg_a = torch.zeros(n_seg, dtype=torch.int64, device=dev).scatter_(0, seg_of_row, row_a)
    seg_qb = torch.zeros(n_seg, dtype=torch.int64, device=dev).scatter_(0, seg_of_row, rows_local // bm)
    cd = torch.cdist(cent, cent, compute_mode="donot_use_mm_for_euclid_dist")
    nonempty = counts > 0
    ts = (starts // bn).to(torch.int64)
    te = torch.where(nonempty, (ends - 1) // bn, ts - 1).to(torch.int64)               # inclusive; empty cluster: no tiles
    keys = []
    chunk = max(1, (1 << 26) // n_centroids)
    for s0 in range(0, n_seg, chunk):
        s1 = min(n_seg, s0 + chunk)
        lb = cd.index_select(0, seg_a[s0:s1]) - seg_rho[s0:s1, None] - radius[None, :]
        need = (lb * (1.0 - SAFETY) <= seg_rk[s0:s1, None] * (1.0 + SAFETY)) & nonempty[None, :]
        sg, cl = need.nonzero(as_tuple=True)
        if sg.numel() == 0:
            continue
        cnt = (te[cl] - ts[cl] + 1)
        rep_qb = torch.repeat_interleave(seg_qb[s0:s1][sg], cnt)
        base = torch.repeat_interleave(ts[cl], cnt)
        offs = torch.arange(rep_qb.numel(), device=dev) - torch.repeat_interleave(torch.cumsum(cnt, 0) - cnt, cnt)
        tl = base + offs
        keep = (tl < t_begin[rep_qb].long()) | (tl >= t_end[rep_qb].long())             # pass 1 already searched these
        keys.append(torch.unique(rep_qb[keep] * n_tiles + tl[keep]))
    qb_ids, tile_ids = [], []
    if keys:
        # a tile can be needed by several segments of a block and by neighbouring clusters: no tile twice per block
        key = torch.unique(torch.cat(keys))
        qb_ids.append(key // n_tiles)
        tile_ids.append(key % n_tiles)
    if qb_ids:
        qb_all = torch.cat(qb_ids)
        tiles_all = torch.cat(tile_ids).to(torch.int32)
    else:
        qb_all = torch.zeros(0, dtype=torch.int64, device=dev)
        tiles_all = torch.zeros(0, dtype=torch.int32, device=dev)
    per_block = torch.bincount(qb_all, minlength=n_qb)
    tile_ptr = torch.zeros(n_qb + 1, dtype=torch.int32, device=dev)
    tile_ptr[1:] = torch.cumsum(per_block, 0).to(torch.int32)
    visited = int(tiles_all.numel()) + int((t_end - t_begin).sum().item())
    fraction = visited / max(1, n_qb * n_tiles)
    last_stats.clear()
    last_stats.update(rows=n, centroids=n_centroids, query_blocks=n_qb, tiles=n_tiles, visited_tile_fraction=fraction,
                      pass1_tiles=int((t_end - t_begin).sum().item()), pass2_tiles=int(tiles_all.numel()))
    mark("bounds and tile lists")
    if fraction > max_fraction:
        return None
    # ---- 6: pass 2
    if tiles_all.numel():
        stage(2, tp=tile_ptr, tl=tiles_all.contiguous(), resume=1)
    mark("pass 2 (listed tiles)")
    # ---- 7: certification + canonical fp32 rescoring, original indices
    stage(4)
    st_host = stats.tolist()
    n_fb = int(st_host[0])
    fb_sorted = fallback[:n_fb].long() + q_lo                                           # positions in the sorted order
    # back to the caller's row order (multi-GPU: gather the ranks' blocks of the sorted order first)
    if world > 1:
        per = D.block_size(n, world)
        idx = D.all_gather_rows(idx, n, per)
        dist = D.all_gather_rows(dist, n, per)
        fb_pad = torch.full((per,), -1, dtype=torch.int64, device=dev)
        fb_pad[:n_fb] = fb_sorted
        import torch.distributed as tdist
        fb_all = torch.empty(world * per, dtype=torch.int64, device=dev)
        tdist.all_gather_into_tensor(fb_all, fb_pad)
        fb_sorted = fb_all[fb_all >= 0]
    out_idx = torch.empty_like(idx)
    out_dist = torch.empty_like(dist)
    out_idx.index_copy_(0, perm, idx)
    out_dist.index_copy_(0, perm, dist)
    last_stats.update(uncertified_rows=int(fb_sorted.numel()), rescored_per_row=st_host[1] / max(st_host[2], 1))
    mark("certify + rescore, gather, un-permute")
    if debug and rank == 0:
        print("  knn_pruned " + ", ".join(f"{b[0]} {(b[1] - a[1]) * 1e3:.1f} ms" for a, b in zip(marks, marks[1:])), flush=True)
    return out_idx, out_dist, perm.index_select(0, fb_sorted)
