"""GPU parity tests for the graph-construction kernels (K1-K6) against oracle/ and the golden
vectors produced by the reference.  Every call goes through the C ABI (ctypes)."""
import os

import numpy as np
import pytest
import torch

from oracle import umap_oracle as orc

pytestmark = pytest.mark.gpu


def _cuda(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _blobs(n, d, centers, seed, spread=5.0):
    rng = np.random.default_rng(seed)
    c = rng.standard_normal((centers, d)) * spread
    lab = rng.integers(0, centers, n)
    return (c[lab] + rng.standard_normal((n, d))).astype(np.float32), lab


def _random_knn_cols(n, k, rng):
    """n x k distinct non-self columns per row, ascending; every third row points at a few hub
    columns so that in-degrees are skewed."""
    cand = rng.integers(0, n, (n, 4 * k))
    hubs = min(n, 64)
    cand[::3, : k // 2] = rng.integers(0, hubs, (cand[::3].shape[0], k // 2))
    col = np.empty((n, k), dtype=np.int32)
    for r in range(n):
        u = np.unique(cand[r][cand[r] != r])
        if u.shape[0] < k:
            u = np.setdiff1d(np.arange(n), [r])
        col[r] = np.sort(rng.permutation(u)[:k])
    return col



# ------------------------------------------------------------------------------- K1 SIMT
@pytest.mark.parametrize("n,d,k", [(700, 37, 15), (300, 64, 30), (130, 5, 7)])
def test_knn_simt_fit_mode_bit_exact(n, d, k):
    from umap_b200 import graph as G
    x, _ = _blobs(n, d, 5, 11)
    x[10] = x[3]
    x[77] = x[3]                       # duplicates: zero distances, ties broken by index
    idx, dist = G.knn_exact_simt(torch.from_numpy(x), torch.from_numpy(x), k, True)
    oi, od = orc.knn_exact(x, x, k, True)
    assert np.array_equal(idx.cpu().numpy(), oi)
    assert np.array_equal(dist.cpu().numpy().view(np.uint32), od.view(np.uint32))


def test_knn_simt_query_mode_and_short_rows():
    from umap_b200 import graph as G
    x, _ = _blobs(200, 24, 4, 12)
    q, _ = _blobs(70, 24, 4, 13)
    idx, dist = G.knn_exact_simt(torch.from_numpy(q), torch.from_numpy(x), 9, False)
    oi, od = orc.knn_exact(q, x, 9, False)
    assert np.array_equal(idx.cpu().numpy(), oi)
    assert np.array_equal(dist.cpu().numpy().view(np.uint32), od.view(np.uint32))
    # fewer admissible points than k: padded with (-1, +inf) exactly like the oracle
    tiny = x[:6]
    idx, dist = G.knn_exact_simt(torch.from_numpy(tiny), torch.from_numpy(tiny), 9, True)
    oi, od = orc.knn_exact(tiny, tiny, 9, True)
    assert np.array_equal(idx.cpu().numpy(), oi)
    assert np.array_equal(dist.cpu().numpy().view(np.uint32), od.view(np.uint32))


def test_knn_simt_streaming_over_db_shards_and_row_subset():
    """db streamed in shards with the running top-k merged in place (the multi-GPU schedule),
    and recomputation of a subset of rows (the fallback for uncertified rows)."""
    from umap_b200 import graph as G
    x, _ = _blobs(500, 16, 5, 14)
    xt = torch.from_numpy(x).cuda()
    k = 15
    out = None
    for lo in range(0, 500, 170):
        hi = min(lo + 170, 500)
        out = G.knn_exact_simt(xt, xt[lo:hi].contiguous(), k, True, query_base=0, db_base=lo, out=out)
    oi, od = orc.knn_exact(x, x, k, True)
    assert np.array_equal(out[0].cpu().numpy(), oi)
    assert np.array_equal(out[1].cpu().numpy().view(np.uint32), od.view(np.uint32))
    # row subset: scribble on some rows, recompute only those
    rows = torch.tensor([3, 77, 128, 499], dtype=torch.int32, device="cuda")
    idx, dist = out[0].clone(), out[1].clone()
    idx[rows.long()] = -7
    G.knn_exact_simt(xt, xt, k, True, out=(idx, dist), rows=rows)
    assert np.array_equal(idx.cpu().numpy(), oi)


def test_knn_merge():
    from umap_b200 import native
    from umap_b200.native import check, lib, ptr, stream
    x, _ = _blobs(300, 12, 3, 15)
    k = 10
    ia, da = orc.knn_exact(x, x[:150], k, True)
    ib, db = orc.knn_exact(x, x[150:], k, True, self_offset=-150)
    ib = np.where(ib >= 0, ib + 150, ib).astype(np.int32)
    oi = torch.empty((300, k), dtype=torch.int32, device="cuda")
    od = torch.empty((300, k), dtype=torch.float32, device="cuda")
    ta, tda, tb, tdb = _cuda(ia), _cuda(da), _cuda(ib), _cuda(db)      # keep the device buffers alive
    check(lib().mmu_knn_merge(ptr(ta), ptr(tda), ptr(tb), ptr(tdb), 300, k, ptr(oi), ptr(od), stream()),
          "mmu_knn_merge")
    ri, rd = orc.knn_exact(x, x, k, True)
    assert np.array_equal(oi.cpu().numpy(), ri)
    assert np.array_equal(od.cpu().numpy().view(np.uint32), rd.view(np.uint32))
    assert native.lib().mmu_abi_version() == 1


# ------------------------------------------------------------------------------- K4
@pytest.mark.parametrize("name", ["sigma_blobs.npz", "sigma_bert.npz"])
def test_smooth_knn_newton_reproduces_reference(golden_dir, name):
    """solver="newton" is the reference's iteration (model.py:33-61) including its divergent
    rows; compared with the reference's own output."""
    from umap_b200 import graph as G
    g = _load(golden_dir, name)
    col, w, sigma, rho = G.smooth_knn(_cuda(g["idx"]), _cuda(g["dist"]), "newton")
    ref = g["sigma"]
    rel = np.abs(sigma.cpu().numpy() - ref) / np.maximum(np.abs(ref), 1e-12)
    assert np.quantile(rel, 0.99) < 1e-4
    assert (rel < 1e-3).mean() > 0.99
    assert np.array_equal(rho.cpu().numpy(), g["dist"].min(axis=1))
    # rows come back ordered by column, as .coalesce() leaves them (model.py:208)
    ci, cw = orc.coalesce_rows(g["idx"], g["weights"])
    assert np.array_equal(col.cpu().numpy(), ci)
    assert np.abs(w.cpu().numpy() - cw).mean() < 1e-5


@pytest.mark.parametrize("name", ["sigma_blobs.npz", "sigma_bert.npz"])
def test_smooth_knn_bisect_matches_oracle_and_reference_where_converged(golden_dir, name):
    from umap_b200 import graph as G
    g = _load(golden_dir, name)
    d = g["dist"]
    k = d.shape[1]
    col, w, sigma, rho = G.smooth_knn(_cuda(g["idx"]), _cuda(d), "bisect")
    sig = sigma.cpu().numpy()
    o = orc.sigmas_bisect(d)
    assert np.allclose(sig, o, rtol=1e-4)
    resid = np.abs(np.exp(-(d - d.min(1, keepdims=True)) / g["sigma"][:, None]).sum(1) - np.log2(k))
    conv = resid < 1e-3                       # rows where the reference's Newton converged
    assert np.allclose(sig[conv], g["sigma"][conv], rtol=2e-4)
    ci, cw = orc.coalesce_rows(g["idx"], orc.membership_weights(d, o))
    assert np.array_equal(col.cpu().numpy(), ci)
    assert np.allclose(w.cpu().numpy(), cw, rtol=1e-3, atol=1e-6)


def test_invert_weights():
    from umap_b200 import graph as G
    x, _ = _blobs(200, 4, 3, 16)
    idx, dist = orc.knn_exact(x, x, 15, False)
    col, w = G.invert_weights(_cuda(idx), _cuda(dist), 1.577, 0.8951)
    ci, cw = orc.coalesce_rows(idx, orc.invert_weights(dist, 1.577, 0.8951))
    assert np.array_equal(col.cpu().numpy(), ci)
    assert np.allclose(w.cpu().numpy(), cw, rtol=1e-5, atol=1e-7)


# ------------------------------------------------------------------------------- K5
def test_fuzzy_union_matches_reference_golden(golden_dir):
    from umap_b200 import graph as G
    g = _load(golden_dir, "union.npz")
    n = int(g["n"])
    k = g["rows"].shape[0] // n
    col = _cuda(g["cols"].reshape(n, k), torch.int32)
    w = _cuda(g["vals"].reshape(n, k))
    s = G.fuzzy_union(col, w)
    assert np.array_equal(s.row.cpu().numpy().astype(np.int64), g["out_rows"])      # bit-exact indices
    assert np.array_equal(s.col.cpu().numpy().astype(np.int64), g["out_cols"])
    rp = s.rowptr.cpu().numpy()
    assert rp[-1] == g["out_rows"].shape[0]
    assert np.array_equal(np.diff(rp), np.bincount(g["out_rows"], minlength=n))
    # values: the oracle's literal fl(fl(a+b)-fl(ab)) bit for bit; <= 2 ulp from torch's
    # sort-order-dependent three-term sum (see tests/test_oracle_golden.py)
    _, _, ov = orc.fuzzy_union(g["rows"], g["cols"], g["vals"], n)
    v = s.val.cpu().numpy()
    assert np.array_equal(v.view(np.uint32), ov.view(np.uint32))
    ulp = np.abs(v.view(np.int32).astype(np.int64) - g["out_vals"].view(np.int32).astype(np.int64))
    assert ulp.max() <= 2


@pytest.mark.parametrize("n,k,seed", [(5000, 15, 1), (70000, 30, 2), (33, 5, 3)])
def test_fuzzy_union_random_graphs(n, k, seed):
    """Multi-pass radix path (n > 2048 bins) and hub columns (skewed in-degree)."""
    from umap_b200 import graph as G
    rng = np.random.default_rng(seed)
    col = _random_knn_cols(n, k, rng)
    w = rng.random((n, k), dtype=np.float32)
    s = G.fuzzy_union(_cuda(col), _cuda(w))
    rows = np.repeat(np.arange(n, dtype=np.int64), k)
    orow, ocol, oval = orc.fuzzy_union(rows, col.reshape(-1).astype(np.int64), w.reshape(-1), n)
    assert np.array_equal(s.row.cpu().numpy().astype(np.int64), orow)
    assert np.array_equal(s.col.cpu().numpy().astype(np.int64), ocol)
    assert np.array_equal(s.val.cpu().numpy().view(np.uint32), oval.view(np.uint32))
    # symmetry: S == S^T as a set of (row, col, val)
    key = orow * n + ocol
    tkey = ocol * n + orow
    assert np.array_equal(np.sort(key), np.sort(tkey))


# ------------------------------------------------------------------------------- K6
def test_embed_query_matches_reference_golden(golden_dir):
    from umap_b200 import graph as G
    g = _load(golden_dir, "embed_query.npz")
    col = _cuda(g["cols"].reshape(50, 15), torch.int32)
    w = _cuda(g["vals"].reshape(50, 15))
    out = G.embed_query(col, w, _cuda(g["ref"]))
    assert np.allclose(out.cpu().numpy(), g["out"], rtol=1e-5, atol=1e-6)


def test_spmm_csr():
    import scipy.sparse as sp
    from umap_b200 import graph as G
    rng = np.random.default_rng(5)
    n, m = 3000, 17
    a = sp.random(n, n, density=0.004, format="csr", dtype=np.float32, random_state=5)
    a.sort_indices()
    coo = a.tocoo()
    t = torch.sparse_coo_tensor(torch.from_numpy(np.stack([coo.row, coo.col]).astype(np.int64)),
                                torch.from_numpy(coo.data), (n, n)).coalesce()
    g = G.Graph.from_sparse_coo(t)
    x = rng.standard_normal((n, m)).astype(np.float32)
    y = G.spmm(g, _cuda(x))
    ref = (a.astype(np.float64) @ x.astype(np.float64))
    assert np.allclose(y.cpu().numpy(), ref, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("m", [32, 17])
def test_spmm_axpby_all_forms(m):
    """y = alpha A x + beta x + gamma z for the 32-column kernel (4 rows per warp) and the generic one:
    ragged rows, empty rows, n not a multiple of 4, z aliasing the output (the Chebyshev recurrence)."""
    import scipy.sparse as sp
    from umap_b200 import graph as G
    rng = np.random.default_rng(m)
    n = 2999
    a = sp.random(n, n, density=0.006, format="lil", dtype=np.float32, random_state=m)
    a[5, :] = 0
    a[n - 1, :] = 0
    a[17, :60] = 1.0                                   # a long row
    a = a.tocsr()
    a.eliminate_zeros()
    a.sort_indices()
    coo = a.tocoo()
    t = torch.sparse_coo_tensor(torch.from_numpy(np.stack([coo.row, coo.col]).astype(np.int64)),
                                torch.from_numpy(coo.data), (n, n)).coalesce()
    g = G.Graph.from_sparse_coo(t)
    x = rng.standard_normal((n, m)).astype(np.float32)
    z = rng.standard_normal((n, m)).astype(np.float32)
    ax = a.astype(np.float64) @ x.astype(np.float64)
    xt, zt = _cuda(x), _cuda(z)
    y = G.spmm_axpby(g, g.val, xt, 1.0, 0.0, None, 0.0)
    assert np.allclose(y.cpu().numpy(), ax, rtol=1e-4, atol=1e-5)
    y = G.spmm_axpby(g, g.val, xt, 0.7, -1.3, None, 0.0)
    assert np.allclose(y.cpu().numpy(), 0.7 * ax - 1.3 * x, rtol=1e-4, atol=1e-5)
    out = zt.clone()
    G.spmm_axpby(g, g.val, xt, 2.0, 0.5, out, -1.0, out=out)       # z aliases the output
    assert np.allclose(out.cpu().numpy(), 2.0 * ax + 0.5 * x - z, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("n,k,blocks", [(5000, 15, 3), (1300, 30, 4), (700, 5, 7)])
def test_union_row_blocks_concatenate_to_the_full_union(n, k, blocks):
    """mmu_fuzzy_union_rows (the per-rank share of the multi-GPU union) on consecutive row blocks, in-edges
    filtered out of the full graph: rowptr, indices and value bits concatenate to exactly what mmu_fuzzy_union gives."""
    from umap_b200 import graph as G
    rng = np.random.default_rng(n + k)
    x = torch.from_numpy(rng.standard_normal((n, 12)).astype(np.float32)).cuda()
    idx, dist = G.knn_graph(x, x, k, True)
    col, w, _, _ = G.smooth_knn(idx, dist, "bisect")
    full = G.fuzzy_union(col, w)
    cuts = [n * i // blocks for i in range(blocks + 1)]
    cuts[1] = max(1, cuts[1] - 37) if blocks > 2 else cuts[1]          # uneven blocks
    rows, cols, vals, counts = [], [], [], []
    for lo, hi in zip(cuts, cuts[1:]):
        rp, r_, c_, v_ = G.fuzzy_union_rows(col, w, lo, hi)
        assert int(rp[0]) == 0 and rp.numel() == hi - lo + 1
        rows.append(r_); cols.append(c_); vals.append(v_); counts.append(rp[1:] - rp[:-1])
    assert torch.equal(torch.cat(rows), full.row) and torch.equal(torch.cat(cols), full.col)
    assert torch.equal(torch.cat(vals).view(torch.int32), full.val.view(torch.int32))
    assert torch.equal(torch.cumsum(torch.cat(counts), 0), full.rowptr[1:])
