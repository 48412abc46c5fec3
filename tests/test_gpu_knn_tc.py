"""GPU parity tests of the tensor-core kNN path (mmu_knn_tc: tcgen05 candidates + certification +
canonical fp32 rescoring, exhaustive fallback for uncertified rows) against the CPU oracle
(oracle/knn_oracle.c, the exhaustive restatement of /root/reference/impl/model.py:109,163,181-193)
and, at sizes the oracle cannot finish in seconds, against the exhaustive CUDA-core kernel.
Bar: indices AND fp32 distance bit patterns identical."""
import numpy as np
import pytest
import torch

from oracle import umap_oracle as orc

pytestmark = pytest.mark.gpu


def _blobs(n, d, centers, seed, spread=5.0, noise=1.0):
    rng = np.random.default_rng(seed)
    c = rng.standard_normal((centers, d)) * spread
    return (c[rng.integers(0, centers, n)] + noise * rng.standard_normal((n, d))).astype(np.float32)


def _tc(q, db, k, excl):
    from umap_b200 import knn_tc
    i, d = knn_tc.knn_tc(q, db, k, excl)
    torch.cuda.synchronize()
    return i.cpu().numpy(), d.cpu().numpy(), dict(knn_tc.last_stats)


def _same(a_idx, a_dist, b_idx, b_dist):
    assert np.array_equal(a_idx, b_idx), f"{(a_idx != b_idx).any(axis=1).sum()} rows differ"
    assert np.array_equal(a_dist.view(np.uint32), b_dist.view(np.uint32))


@pytest.mark.parametrize("n,d,k", [(700, 37, 15), (3000, 64, 15), (1500, 200, 30), (513, 768, 5), (130, 5, 7)])
def test_fit_mode_bit_exact_vs_oracle(n, d, k):
    x = _blobs(n, d, 6, n + d)
    xt = torch.from_numpy(x).cuda()
    ti, td, st = _tc(xt, xt, k, True)
    oi, od = orc.knn_exact(x, x, k, True)
    _same(ti, td, oi, od)
    assert st["rows"] == n


def test_query_mode_bit_exact_vs_oracle():
    db = _blobs(2100, 96, 5, 1)
    q = _blobs(333, 96, 5, 2)
    ti, td, _ = _tc(torch.from_numpy(q).cuda(), torch.from_numpy(db).cuda(), 15, False)
    oi, od = orc.knn_exact(q, db, 15, False)
    _same(ti, td, oi, od)


def test_uncentred_large_offset_data():
    """A large common offset (BERT-pooler-like bias) must not cost exactness: the prep centres it."""
    x = _blobs(2000, 128, 8, 3, spread=1.0) + np.float32(40.0)
    xt = torch.from_numpy(x).cuda()
    ti, td, st = _tc(xt, xt, 15, True)
    oi, od = orc.knn_exact(x, x, 15, True)
    _same(ti, td, oi, od)
    assert st["fallback_rows"] < 0.05 * st["rows"], st


def test_duplicates_and_exact_ties_fall_back_correctly():
    """Many exact duplicates: distance ties broken by index; rows the bound cannot certify must
    be finished by the exhaustive kernel and still match the oracle bit for bit."""
    rng = np.random.default_rng(5)
    base = rng.standard_normal((40, 24)).astype(np.float32)
    x = base[rng.integers(0, 40, 1200)]                     # every row has ~30 exact copies
    xt = torch.from_numpy(x).cuda()
    ti, td, st = _tc(xt, xt, 15, True)
    oi, od = orc.knn_exact(x, x, 15, True)
    _same(ti, td, oi, od)
    # >64 identical copies of each row exceed the candidate list: certification must fail there
    y = base[rng.integers(0, 8, 1500)]
    yt = torch.from_numpy(y).cuda()
    ti, td, st = _tc(yt, yt, 15, True)
    oi, od = orc.knn_exact(y, y, 15, True)
    _same(ti, td, oi, od)
    assert st["fallback_rows"] > 0


def test_constant_and_tiny_inputs():
    x = np.zeros((300, 16), np.float32)
    xt = torch.from_numpy(x).cuda()
    ti, td, _ = _tc(xt, xt, 5, True)
    oi, od = orc.knn_exact(x, x, 5, True)
    _same(ti, td, oi, od)
    x = _blobs(40, 8, 2, 9)
    xt = torch.from_numpy(x).cuda()
    ti, td, _ = _tc(xt, xt, 10, True)
    oi, od = orc.knn_exact(x, x, 10, True)
    _same(ti, td, oi, od)


@pytest.mark.parametrize("n,d,kind", [(20000, 256, "blobs"), (12000, 768, "bert"), (6000, 4096, "vae")])
def test_matches_exhaustive_cuda_kernel_at_scale(n, d, kind):
    """Shapes of BASELINE.json configs[1] (BERT 768-D, SD-VAE 4096-D): the tensor-core path must
    equal the exhaustive fp32 kernel everywhere and certify nearly every row itself."""
    from umap_b200 import graph as G
    g = torch.Generator(device="cuda").manual_seed(n)
    cl = torch.arange(n, device="cuda") % 64
    if kind == "bert":
        x = torch.tanh(torch.randn((64, d), generator=g, device="cuda")[cl] + 0.5 * torch.randn((n, d), generator=g, device="cuda"))
    elif kind == "vae":
        x = 2.0 * torch.randn((64, d), generator=g, device="cuda")[cl] + 4.0 * torch.randn((n, d), generator=g, device="cuda")
    else:
        x = 5.0 * torch.randn((64, d), generator=g, device="cuda")[cl] + torch.randn((n, d), generator=g, device="cuda")
    x = x.contiguous()
    ti, td, st = _tc(x, x, 15, True)
    si, sd = G.knn_exact_simt(x, x, 15, True)
    _same(ti, td, si.cpu().numpy(), sd.cpu().numpy())
    assert st["fallback_rows"] <= 0.02 * n, st


def test_database_ring_emulated_on_one_gpu():
    """The row-sharded-database search (umap_b200.dist.ring_knn's per-shard step): searching the
    database shard by shard with mmu_knn_tc and merging with mmu_knn_merge equals the one-shot
    search, including self-exclusion in the shard that holds the query rows."""
    from umap_b200 import knn_tc
    from umap_b200.native import check, lib, ptr, stream
    x = _blobs(3000, 64, 6, 21)
    xt = torch.from_numpy(x).cuda()
    k, shards = 15, [(0, 1024), (1024, 2048), (2048, 3000)]
    q_lo, q_hi = 1024, 2048
    q = xt[q_lo:q_hi].contiguous()
    best = None
    for lo, hi in shards:
        same = lo == q_lo
        i_s, d_s = knn_tc.knn_tc(q, xt[lo:hi].contiguous(), k, same, query_base=0)
        i_s = torch.where(i_s >= 0, i_s + lo, i_s)
        if best is None:
            best = (i_s, d_s)
        else:
            oi = torch.empty_like(i_s)
            od = torch.empty_like(d_s)
            check(lib().mmu_knn_merge(ptr(best[0]), ptr(best[1]), ptr(i_s), ptr(d_s), q.shape[0], k, ptr(oi), ptr(od),
                                      stream()), "merge")
            best = (oi, od)
    oi, od = orc.knn_exact(x, x, k, True)
    _same(best[0].cpu().numpy(), best[1].cpu().numpy(), oi[q_lo:q_hi], od[q_lo:q_hi])


def test_deep_candidate_pool_certifies_what_one_list_cannot():
    """Low dimension, k=30, dense neighbourhoods: the gap between the 30th and the 64th neighbour is
    below the fp16 error bound, so the first pass leaves rows uncertified; the retry with the
    database in 8 splits (512 candidates per row) and error-compensated split-fp16 operands must
    certify them -- and the result must still equal the exhaustive kernel bit for bit."""
    from umap_b200 import graph as G
    g = torch.Generator(device="cuda").manual_seed(4)
    n, d, k = 60000, 16, 30
    x = (5.0 * torch.randn((20, d), generator=g, device="cuda")[torch.arange(n, device="cuda") % 20]
         + torch.randn((n, d), generator=g, device="cuda")).contiguous()
    ti, td, st = _tc(x, x, k, True)
    si, sd = G.knn_exact_simt(x, x, k, True)
    _same(ti, td, si.cpu().numpy(), sd.cpu().numpy())
    assert st["first_pass_uncertified"] > 0.02 * n or st["precision"] == 1, st      # the fp16 pass alone is not enough here
    assert st["fallback_rows"] < 0.001 * n, st                                       # ... the split-fp16 deep pool is


@pytest.mark.parametrize("n,d,kind", [(12000, 768, "bert"), (5000, 4096, "vae")])
def test_cta_pairs_windowed_and_single_cta_forms_agree(n, d, kind):
    """The candidate kernel has two forms (knn_tc.cu: cta_group::2 CTA pairs for long rows, one CTA per
    query block otherwise) and the pair form walks a large database in windows, carrying the per-row
    lists from launch to launch.  All of them must give the exhaustive kernel's result bit for bit;
    an odd number of query blocks (the last pair is half padding) is part of the case."""
    from umap_b200 import graph as G
    g = torch.Generator(device="cuda").manual_seed(n + 1)
    cl = torch.arange(n, device="cuda") % 64
    if kind == "bert":
        x = torch.tanh(torch.randn((64, d), generator=g, device="cuda")[cl] + 0.5 * torch.randn((n, d), generator=g, device="cuda"))
    else:
        x = 2.0 * torch.randn((64, d), generator=g, device="cuda")[cl] + 4.0 * torch.randn((n, d), generator=g, device="cuda")
    x = x.contiguous()
    si, sd = G.knn_exact_simt(x, x, 15, True)
    si, sd = si.cpu().numpy(), sd.cpu().numpy()
    q = x[: 128 * 37 + 5].contiguous()                       # 38 query blocks in query mode, 37 full
    qi, qd = G.knn_exact_simt(q, x, 15, False)
    qi, qd = qi.cpu().numpy(), qd.cpu().numpy()
    from umap_b200 import native
    defaults = {"knn_window_mb": native.get_option("knn_window_mb"), "knn_cta_pairs": native.get_option("knn_cta_pairs")}
    try:
        for opts in ({"knn_window_mb": 1}, {"knn_window_mb": 0}, {"knn_cta_pairs": 0}):
            for key, val in opts.items():
                native.set_option(key, val)
            ti, td, st = _tc(x, x, 15, True)
            _same(ti, td, si, sd)
            assert st["fallback_rows"] <= 0.02 * n, (opts, st)
            ti, td, st = _tc(q, x, 15, False)
            _same(ti, td, qi, qd)
            for key in opts:
                native.set_option(key, defaults[key])
    finally:
        for key, val in defaults.items():
            native.set_option(key, val)


@pytest.mark.parametrize("n,d,k,blobs,cents", [(60000, 64, 30, 200, 256), (40000, 128, 15, 50, 64), (30011, 16, 10, 300, 320)])
def test_cluster_pruned_search_bit_exact(n, d, k, blobs, cents):
    """knn_pruned.knn_pruned (rows sorted by cluster, per-block tile ranges / lists, ball bounds, original indices
    through db_gid) must equal the exhaustive kernel bit for bit on every row it reports as done, and must actually
    prune on clustered data; the rows it hands back as uncertified are finished by knn_tc's deeper levels."""
    from umap_b200 import graph as G
    from umap_b200 import knn_pruned, knn_tc
    g = torch.Generator(device="cuda").manual_seed(n + d)
    centres = 5.0 * torch.randn((blobs, d), generator=g, device="cuda")
    x = (centres[torch.arange(n, device="cuda") % blobs] + torch.randn((n, d), generator=g, device="cuda")).contiguous()
    x[123] = x[77]                                            # an exact duplicate: a distance tie decided by ORIGINAL index
    x[5000:5040] = x[5000]                                    # and a run of 40 identical rows
    assert knn_pruned.contrast(x, k) < knn_tc.PRUNE_CONTRAST
    res = knn_pruned.knn_pruned(x, k, n_centroids=cents)
    assert res is not None, knn_pruned.last_stats
    idx, dist, fb = res
    st = dict(knn_pruned.last_stats)
    assert st["visited_tile_fraction"] < 0.5, st             # (clusters far smaller than a 256-row tile prune poorly)
    si, sd = G.knn_exact_simt(x, x, k, True)
    done = torch.ones(n, dtype=torch.bool, device="cuda")
    done[fb] = False
    assert int(fb.numel()) < 0.01 * n, st
    assert torch.equal(idx[done], si[done]), f"{int((idx[done] != si[done]).any(dim=1).sum())} rows differ"
    assert torch.equal(dist[done].view(torch.int32), sd[done].view(torch.int32))


def test_pruned_search_declines_on_unclustered_data():
    """Uniform noise has no cluster structure: the contrast heuristic says so, and a forced attempt finds that its tile
    lists cover most of the database and returns None (the caller then runs the full contraction)."""
    from umap_b200 import knn_pruned, knn_tc
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((40000, 64), generator=g, device="cuda").contiguous()
    assert knn_pruned.contrast(x, 15) > knn_tc.PRUNE_CONTRAST
    assert knn_pruned.knn_pruned(x, 15, n_centroids=256) is None
    assert knn_pruned.last_stats["visited_tile_fraction"] > 0.5


def test_farthest_point_kernel_covers_every_cluster():
    """mmu_fps_centroids (persistent kernel, grid barrier per round): the first centroids of well separated blobs are
    one per blob, and the selection equals the torch restatement of the greedy rule wherever the arg-max is not a
    floating-point near-tie (the two forms sum the squared distances in different orders)."""
    from umap_b200 import knn_pruned
    g = torch.Generator(device="cuda").manual_seed(4)
    blobs, d, n = 37, 48, 20000
    centres = 6.0 * torch.randn((blobs, d), generator=g, device="cuda")
    lab = torch.arange(n, device="cuda") % blobs
    x = (centres[lab] + torch.randn((n, d), generator=g, device="cuda")).contiguous()
    cent = knn_pruned.farthest_point_centroids(x, 64)
    ref = knn_pruned.farthest_point_centroids_torch(x, 64)
    torch.cuda.synchronize()
    near = torch.cdist(cent[:blobs], centres).argmin(dim=1)
    assert near.unique().numel() == blobs                       # one centroid in every blob before any blob gets two
    same = (cent == ref).all(dim=1)
    assert int(same[:blobs].sum()) >= blobs - 2 and bool(same[0])
    assert bool(torch.isfinite(cent).all())
