"""GPU parity tests for the layout optimiser kernels (K7 forces, K8 InfoNCE, K9 Adam) against
the reference's golden vectors and the oracle, driven by the reference's own host sample stream
(torch CPU generator replay).  Tolerances are stated per test; their floor is the reference's
own rerun spread."""
import os

import numpy as np
import pytest
import torch

from oracle import umap_oracle as orc

pytestmark = pytest.mark.gpu


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _coo(rows, cols, vals, shape):
    idx = torch.from_numpy(np.stack([rows, cols]).astype(np.int64))
    return torch.sparse_coo_tensor(idx, torch.from_numpy(vals.astype(np.float32)), shape).coalesce()


def _opt(embeds, graphs, g, mode="fit", refs=None, stream="host", **kw):
    from umap_b200.layout import LayoutOptimizer
    return LayoutOptimizer([torch.from_numpy(e) for e in embeds], graphs, float(g["a"]), float(g["b"]),
                           int(g["num_rep"]), float(g["lr"]), float(g["alpha"]) if "alpha" in g else 1.0,
                           int(g["batch_size"]), mode=mode, refs=refs, sample_stream=stream, **kw)


@pytest.mark.parametrize("epochs,tol", [(1, 5e-6), (5, 5e-5), (20, 1e-3)])
def test_fit_matches_reference_golden(golden_dir, epochs, tol):
    """Final embeddings of UMAPMixture._train (model.py:396-481) after `epochs` epochs under the
    same torch.manual_seed: max abs difference < tol (1 epoch: Adam's first step is +-lr, so
    this pins the SIGN and support of every gradient entry; longer horizons amplify fp32
    rounding chaotically -- the reference differs from its own rerun by 7e-6 at 10 epochs)."""
    g = _load(golden_dir, "train_fit.npz")
    graphs = [_coo(g[f"rows{m}"], g[f"cols{m}"], g[f"vals{m}"], (g[f"init{m}"].shape[0],) * 2) for m in range(2)]
    torch.manual_seed(int(g["seed"]))
    opt = _opt([g["init0"], g["init1"]], graphs, g)
    out = opt.run(epochs)
    for m in range(2):
        diff = np.abs(out[m].cpu().numpy() - g[f"fit{epochs}_{m}"])
        assert diff.max() < tol, (m, diff.max())


@pytest.mark.parametrize("epochs,tol", [(1, 5e-6), (10, 1e-4)])
def test_transform_matches_reference_golden(golden_dir, epochs, tol):
    g = _load(golden_dir, "train_transform.npz")
    graph = _coo(g["rows"], g["cols"], g["vals"], (g["init"].shape[0], g["ref"].shape[0]))
    torch.manual_seed(int(g["seed"]))
    opt = _opt([g["init"]], [graph], g, mode="transform", refs=[torch.from_numpy(g["ref"])])
    out = opt.run(epochs)
    diff = np.abs(out[0].cpu().numpy() - g[f"tr{epochs}"])
    assert diff.max() < tol, diff.max()


@pytest.mark.parametrize("fast", [False, True])
@pytest.mark.parametrize("num_rep", [5, 8, 4])
@pytest.mark.parametrize("dim", [2, 4, 8, 16, 32, 64, 128, 3, 20])
def test_single_epoch_gradient_matches_oracle(dim, num_rep, fast):
    """One epoch from zero Adam state moves every coordinate by -lr*g/(|g|+eps'): compare the
    gradient buffer itself (before Adam) with the fp64 closed form of the oracle, all vectorised
    kernel variants (dim = 2..128), the generic one, the loop version (num_rep=5) and the
    register-blocked version (num_rep 4/8); fast = ex2/lg2/rcp arithmetic of the device stream
    (tolerance 1e-4 of the largest gradient entry instead of 2e-5)."""
    from umap_b200.layout import LayoutOptimizer, replay_host_draws
    from umap_b200.native import check, lib, ptr, stream
    rng = np.random.default_rng(dim)
    n, k, bs = 400, 10, 128
    y = (rng.standard_normal((n, dim)) * 0.5).astype(np.float32)
    y[7] = y[3]                                       # zero distance -> clamp branch, zero gradient
    cols = np.stack([np.sort(rng.choice(np.setdiff1d(np.arange(n), [r]), k, replace=False)) for r in range(n)])
    cols[3, 0] = 7 if 7 not in cols[3] else cols[3, 0]
    cols[3] = np.sort(cols[3])
    rows = np.repeat(np.arange(n), k)
    vals = rng.random(n * k).astype(np.float32)
    graph = _coo(rows, cols.reshape(-1), vals, (n, n))
    gi = graph.indices().numpy()
    gv = graph.values().numpy()
    a, b = 1.577, 0.8951
    opt = LayoutOptimizer([torch.from_numpy(y)], [graph], a, b, num_rep, 0.01, 1.0, bs, mode="fit",
                          sample_stream="host", track_loss=True)
    opt.fast_math = fast
    mod = opt.mods[0]
    torch.manual_seed(123)
    kept, neg, counts = replay_host_draws(mod, num_rep)
    nk = mod.load_host_draws(kept, counts)
    neg_d = neg.cuda()
    opt._forces(mod, mod.kept_rec, mod.kept_hdr, neg_d, mod.batch_kept)
    torch.cuda.synchronize()
    got = mod.g.cpu().numpy().astype(np.float64)
    # oracle: same draws
    y64 = y.astype(np.float64)
    grad = np.zeros_like(y64)
    nb = (n + bs - 1) // bs
    kept_np, neg_np = kept.numpy(), neg.numpy()
    off = 0
    loss = 0.0
    for bi in range(nb):
        c = int(counts[bi])
        pos = kept_np[off:off + c]
        ii, jj = gi[0][pos], gi[1][pos]
        la, ga = orc.umap_attr_grad(y64[ii], y64[jj], a, b)
        np.add.at(grad, ii, ga / nb)
        np.add.at(grad, jj, -ga / nb)
        ll = neg_np[off:off + c].reshape(-1)
        ir = np.repeat(ii, num_rep)
        lr_, gr = orc.umap_rep_grad(y64[ir], y64[ll], a, b)
        np.add.at(grad, ir, gr / nb)
        np.add.at(grad, ll, -gr / nb)
        loss += (la + lr_) / nb
        off += c
    scale = np.abs(grad).max()
    assert np.abs(got - grad).max() < (1e-4 if fast else 2e-5) * scale + 1e-9
    assert abs(float(opt.loss.item()) - loss) < 1e-4 * abs(loss)
    assert gv.shape[0] == n * k


def test_infonce_matches_reference_autograd(golden_dir):
    from umap_b200.layout import replay_infonce_draws
    from umap_b200.native import check, lib, ptr, stream
    g = _load(golden_dir, "losses.npz")
    e0, e1 = torch.from_numpy(g["e0"]).cuda(), torch.from_numpy(g["e1"]).cuda()
    num = min(e0.shape[0], e1.shape[0])
    torch.manual_seed(int(g["infonce_seed"]))
    perm, neg = replay_infonce_draws(num)
    g0, g1 = torch.zeros_like(e0), torch.zeros_like(e1)
    loss = torch.zeros(1, device="cuda")
    state = torch.zeros(8, dtype=torch.int32, device="cuda")
    perm_d, neg_d = perm.cuda(), neg.cuda()
    check(lib().mmu_infonce(ptr(e0), ptr(e1), num, e0.shape[1], ptr(perm_d), ptr(neg_d), 9, 1000, 1.0, 0.5,
                            ptr(g0), ptr(g1), 0, 0, ptr(state), ptr(loss), stream()), "mmu_infonce")
    assert abs(loss.item() - float(g["infonce_loss"])) < 1e-4
    s0, s1 = np.abs(g["infonce_g0"]).max(), np.abs(g["infonce_g1"]).max()
    assert np.abs(g0.cpu().numpy() - g["infonce_g0"]).max() < 1e-4 * s0
    assert np.abs(g1.cpu().numpy() - g["infonce_g1"]).max() < 1e-4 * s1


@pytest.mark.parametrize("dim", [4, 8, 16, 32, 64, 128, 5])
def test_infonce_all_kernel_variants_match_oracle(dim):
    """Vectorised (dim = 4..128) and generic InfoNCE kernels vs the fp64 closed form of the oracle
    (model.py:364-394), including masked negatives (model.py:386) and a ragged last chunk."""
    from umap_b200.native import check, lib, ptr, stream
    rng = np.random.default_rng(dim)
    n0, n1, num = 2600, 2300, 2300
    e0 = rng.standard_normal((n0, dim)).astype(np.float32)
    e1 = rng.standard_normal((n1, dim)).astype(np.float32)
    perm = rng.permutation(num)
    negs = rng.integers(0, num, (num, 9))
    negs[::5, 2] = perm[::5]
    l_ref, g0_ref, g1_ref = orc.infonce_grad_vec(e0, e1, perm, negs)
    e0t, e1t = torch.from_numpy(e0).cuda(), torch.from_numpy(e1).cuda()
    g0, g1 = torch.zeros_like(e0t), torch.zeros_like(e1t)
    loss = torch.zeros(1, device="cuda")
    state = torch.zeros(8, dtype=torch.int32, device="cuda")
    perm_d = torch.from_numpy(perm.astype(np.int32)).cuda()         # keep alive across the async launch
    negs_d = torch.from_numpy(negs.astype(np.int32)).cuda()
    check(lib().mmu_infonce(ptr(e0t), ptr(e1t), num, dim, ptr(perm_d), ptr(negs_d), 9, 1000, 1.0, 0.5, ptr(g0), ptr(g1),
                            0, 0, ptr(state), ptr(loss), stream()), "mmu_infonce")
    assert abs(loss.item() - l_ref) < 1e-4 * abs(l_ref)
    assert np.abs(g0.cpu().numpy() - g0_ref).max() < 1e-4 * np.abs(g0_ref).max()
    assert np.abs(g1.cpu().numpy() - g1_ref).max() < 1e-4 * np.abs(g1_ref).max()


def test_adam_matches_oracle_bitwise_over_steps():
    from umap_b200.native import check, lib, ptr, stream
    rng = np.random.default_rng(9)
    n = 5000
    p = rng.standard_normal(n).astype(np.float32)
    m = np.zeros(n, np.float32)
    v = np.zeros(n, np.float32)
    pt, mt, vt = (torch.from_numpy(a.copy()).cuda() for a in (p, m, v))
    state = torch.zeros(8, dtype=torch.int32, device="cuda")
    check(lib().mmu_opt_state_init(ptr(state), stream()), "init")
    for step in range(1, 6):
        gnp = (rng.standard_normal(n) * 10.0 ** rng.integers(-6, 2, n)).astype(np.float32)
        gt = torch.from_numpy(gnp.copy()).cuda()
        check(lib().mmu_opt_state_advance(ptr(state), 0.01, 0.9, 0.999, stream()), "advance")
        check(lib().mmu_adam_step(ptr(pt), ptr(gt), ptr(mt), ptr(vt), n, 0.9, 0.999, 1e-8, ptr(state), 1, stream()),
              "adam")
        p, m, v = orc.adam_step(p, gnp, m, v, step, 0.01)
        assert float(gt.abs().max()) == 0.0                     # zero_grad fused
        # same operations, one rounding each (no fma contraction in the kernel): moments bit-exact;
        # the parameter allows 2 ulp for a last-bit difference of the device pow() in step_size
        assert np.array_equal(mt.cpu().numpy().view(np.uint32), m.view(np.uint32))
        assert np.array_equal(vt.cpu().numpy().view(np.uint32), v.view(np.uint32))
        assert np.allclose(pt.cpu().numpy(), p, rtol=2.4e-7, atol=1e-9)
    # against torch.optim.Adam itself
    q = torch.nn.Parameter(torch.from_numpy(rng.standard_normal(64).astype(np.float32)))
    q0 = q.detach().clone().cuda()
    optim = torch.optim.Adam([q], lr=0.01)
    mm, vv = torch.zeros_like(q0), torch.zeros_like(q0)
    check(lib().mmu_opt_state_init(ptr(state), stream()), "init")
    for step in range(3):
        gr = torch.from_numpy(rng.standard_normal(64).astype(np.float32))
        q.grad = gr.clone()
        optim.step()
        gd = gr.cuda()
        check(lib().mmu_opt_state_advance(ptr(state), 0.01, 0.9, 0.999, stream()), "advance")
        check(lib().mmu_adam_step(ptr(q0), ptr(gd), ptr(mm), ptr(vv), 64, 0.9, 0.999, 1e-8, ptr(state), 0, stream()),
              "adam")
        assert np.allclose(q0.cpu().numpy(), q.detach().numpy(), rtol=1e-6, atol=1e-9)


def test_device_stream_statistics_and_determinism():
    """Throughput mode: Philox Bernoulli(w) keeps ~sum(w) edges (model.py:432), per-batch counts
    add up, the same seed reproduces the same kept set, epochs differ."""
    from umap_b200.native import check, lib, ptr, stream
    rng = np.random.default_rng(3)
    n, k, bs = 20000, 15, 256
    rows = np.repeat(np.arange(n, dtype=np.int32), k)
    cols = rng.integers(0, n, n * k).astype(np.int32)
    w = rng.random(n * k).astype(np.float32)
    w[:100] = 1.0
    w[100:200] = 0.0
    row_t, col_t, w_t = torch.from_numpy(rows).cuda(), torch.from_numpy(cols).cuda(), torch.from_numpy(w).cuda()
    nb = (n + bs - 1) // bs
    state = torch.zeros(8, dtype=torch.int32, device="cuda")
    check(lib().mmu_opt_state_init(ptr(state), stream()), "init")

    def sample(seed, cap=n * k):
        rec = torch.full((cap, 4), -1, dtype=torch.int32, device="cuda")
        hdr = torch.tensor([0, cap, 0, 0], dtype=torch.int32, device="cuda")
        bk = torch.zeros(nb, dtype=torch.int32, device="cuda")
        check(lib().mmu_edge_sample_range(ptr(row_t), ptr(col_t), ptr(w_t), 0, n * k, bs, nb, seed, ptr(state), ptr(rec),
                                          ptr(hdr), ptr(bk), stream()), "sample")
        c, _, over, _ = hdr.tolist()
        r = rec[: min(c, cap)].cpu().numpy()
        if not over:
            # a record is {edge position, row, col, row-batch} of a kept edge
            assert np.array_equal(r[:, 1], rows[r[:, 0]]) and np.array_equal(r[:, 2], cols[r[:, 0]])
            assert np.array_equal(r[:, 3], r[:, 1] // bs)
            # row sorted inside every 1024-edge chunk of the sampler
            chunk = r[:, 0] // 1024
            same = chunk[1:] == chunk[:-1]
            assert np.all(r[1:, 0][same] > r[:-1, 0][same])
        return (np.sort(r[:, 0]), bk.cpu().numpy()) if cap == n * k else (c, over)

    k1, b1 = sample(42)
    k2, b2 = sample(42)
    k3, _ = sample(43)
    assert np.array_equal(k1, k2) and np.array_equal(b1, b2)
    assert not np.array_equal(k1, k3)
    assert b1.sum() == k1.shape[0]
    assert np.array_equal(np.bincount(rows[k1] // bs, minlength=nb), b1)
    exp = w.sum()
    assert abs(k1.shape[0] - exp) < 5 * np.sqrt((w * (1 - w)).sum())
    assert np.all(np.isin(np.arange(100), k1)) and not np.any(np.isin(np.arange(100, 200), k1))
    check(lib().mmu_opt_state_advance(ptr(state), 0.01, 0.9, 0.999, stream()), "advance")
    k4, _ = sample(42)
    assert not np.array_equal(k1, k4)
    # a record list that is too small: the surplus is dropped and the overflow flag is raised
    c, over = sample(42, cap=1000)
    assert c > 1000 and over == 1


def test_device_stream_fit_reaches_reference_quality(golden_dir):
    """Device-stream run vs host-stream run of the same problem: same loss trajectory within
    sampling noise (statistical parity of the two sample streams)."""
    g = _load(golden_dir, "train_fit.npz")
    graphs = [_coo(g[f"rows{m}"], g[f"cols{m}"], g[f"vals{m}"], (g[f"init{m}"].shape[0],) * 2) for m in range(2)]
    torch.manual_seed(1)
    oh = _opt([g["init0"], g["init1"]], graphs, g, stream="host", track_loss=True)
    oh.run(30)
    od = _opt([g["init0"], g["init1"]], graphs, g, stream="device", seed=7, track_loss=True)
    od.run(30)
    lh, ld = np.array(oh.losses), np.array(od.losses)
    assert np.all(np.isfinite(ld))
    assert abs(lh[-10:].mean() - ld[-10:].mean()) < 0.05 * abs(lh[-10:].mean())


def test_edge_and_anchor_ranges_partition_the_single_gpu_stream():
    """Multi-GPU sharding contract on one device: sampling edge ranges [0,a),[a,b),[b,nnz) keeps
    exactly the edges the full-range call keeps (Philox keyed on global positions), and the InfoNCE
    gradients of anchor ranges add up to the full-range gradient."""
    from umap_b200.native import check, lib, ptr, stream
    rng = np.random.default_rng(11)
    n, k, bs = 5000, 15, 256
    rows = torch.from_numpy(np.repeat(np.arange(n, dtype=np.int32), k)).cuda()
    cols = torch.from_numpy(rng.integers(0, n, n * k).astype(np.int32)).cuda()
    w = torch.from_numpy(rng.random(n * k).astype(np.float32)).cuda()
    nnz, nb = n * k, (n + bs - 1) // bs
    state = torch.zeros(8, dtype=torch.int32, device="cuda")
    check(lib().mmu_opt_state_init(ptr(state), stream()), "init")

    def sample(lo, hi):
        rec = torch.empty((nnz, 4), dtype=torch.int32, device="cuda")
        hdr = torch.tensor([0, nnz, 0, 0], dtype=torch.int32, device="cuda")
        bk = torch.zeros(nb, dtype=torch.int32, device="cuda")
        check(lib().mmu_edge_sample_range(ptr(rows), ptr(cols), ptr(w), lo, hi, bs, nb, 99, ptr(state), ptr(rec), ptr(hdr),
                                          ptr(bk), stream()), "sample_range")
        return np.sort(rec[: int(hdr[0].item()), 0].cpu().numpy()), bk.cpu().numpy()

    full, bk_full = sample(0, nnz)
    cuts = [0, 1001, 40007, nnz]                      # deliberately not multiples of 4
    parts = [sample(a, b) for a, b in zip(cuts, cuts[1:])]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), full)
    assert np.array_equal(sum(p[1] for p in parts), bk_full)
    for (a, b), (kp, _) in zip(zip(cuts, cuts[1:]), parts):
        assert kp.size == 0 or (kp.min() >= a and kp.max() < b)

    dim, num = 16, 3300
    e0 = torch.from_numpy(rng.standard_normal((num, dim)).astype(np.float32)).cuda()
    e1 = torch.from_numpy(rng.standard_normal((num + 50, dim)).astype(np.float32)).cuda()

    def nce(lo, hi):
        g0, g1 = torch.zeros_like(e0), torch.zeros_like(e1)
        loss = torch.zeros(1, device="cuda")
        check(lib().mmu_infonce_range(ptr(e0), ptr(e1), num, lo, hi, dim, None, None, 9, 1000, 1.0, 0.5, ptr(g0), ptr(g1),
                                      7, 0, ptr(state), ptr(loss), stream()), "infonce_range")
        return g0.cpu().numpy().astype(np.float64), g1.cpu().numpy().astype(np.float64), float(loss.item())

    f0, f1, fl = nce(0, num)
    a0, a1, al = nce(0, 1234)
    b0, b1, bl = nce(1234, num)
    assert np.abs(a0 + b0 - f0).max() < 1e-6 * np.abs(f0).max()
    assert np.abs(a1 + b1 - f1).max() < 2e-6 * np.abs(f1).max()
    assert abs(al + bl - fl) < 1e-5 * abs(fl)


@pytest.mark.parametrize("dim,num_rep", [(24, 4), (130, 8), (4096, 8)])
def test_invert_forces_gradient_matches_oracle(dim, num_rep):
    """mmu_invert_forces vs the fp64 closed form of oracle.inv_attr_grad / inv_rep_grad (pinned to the
    reference's _inv_attr_loss/_inv_rep_loss autograd by tests/golden/invert_losses.npz), with the
    host sample stream; includes a zero-distance pair (clamp) and a rho larger than the distance."""
    from umap_b200.layout import LayoutOptimizer, replay_host_draws
    rng = np.random.default_rng(dim)
    n_ref, q, k, bs = 300, 90, 8, 32
    data = (rng.standard_normal((n_ref, dim)) * 1.2).astype(np.float32)
    cols = np.stack([np.sort(rng.choice(n_ref, k, replace=False)) for _ in range(q)])
    x = (data[cols[:, 0]] + 0.3 * rng.standard_normal((q, dim))).astype(np.float32)
    x[5] = data[cols[5, 0]]
    sigma = (rng.random(n_ref) * 2.0 + 0.1).astype(np.float32)
    rho = (rng.random(n_ref) * np.sqrt(dim)).astype(np.float32)
    rho[cols[6, 1]] = 1e4
    rows = np.repeat(np.arange(q), k)
    vals = (0.3 + 0.7 * rng.random(q * k)).astype(np.float32)
    graph = _coo(rows, cols.reshape(-1), vals, (q, n_ref))
    gi = graph.indices().numpy()
    a, b = 1.577, 0.8951
    opt = LayoutOptimizer([torch.from_numpy(x)], [graph], a, b, num_rep, 0.01, 1.0, bs, mode="invert",
                          refs=[torch.from_numpy(data)], sigmas=[torch.from_numpy(sigma)], rhos=[torch.from_numpy(rho)],
                          sample_stream="host", track_loss=True)
    mod = opt.mods[0]
    torch.manual_seed(321)
    kept, neg, counts = replay_host_draws(mod, num_rep)
    nk = mod.load_host_draws(kept, counts)
    neg_d = neg.cuda()
    opt._forces(mod, mod.kept_rec, mod.kept_hdr, neg_d, mod.batch_kept)
    torch.cuda.synchronize()
    got = mod.g.cpu().numpy().astype(np.float64)
    x64, d64, s64, r64 = x.astype(np.float64), data.astype(np.float64), sigma.astype(np.float64), rho.astype(np.float64)
    grad = np.zeros_like(x64)
    nb = (q + bs - 1) // bs
    kept_np, neg_np = kept.numpy(), neg.numpy()
    off, loss = 0, 0.0
    for bi in range(nb):
        c = int(counts[bi])
        pos = kept_np[off:off + c]
        ii, jj = gi[0][pos], gi[1][pos]
        la, ga = orc.inv_attr_grad(x64[ii], d64[jj], s64[jj], a, b)
        np.add.at(grad, ii, ga / nb)
        ll = neg_np[off:off + c].reshape(-1)
        ir = np.repeat(ii, num_rep)
        lr_, gr = orc.inv_rep_grad(x64[ir], d64[ll], s64[ll], r64[ll])
        np.add.at(grad, ir, gr / nb)
        loss += (la + lr_) / nb
        off += c
    assert np.abs(got - grad).max() < 1e-4 * np.abs(grad).max() + 1e-9
    assert abs(float(opt.loss.item()) - loss) < 1e-3 * abs(loss)


def _device_epoch_gradient(opt):
    """one device-stream epoch's gradient buffer (sample + forces, no Adam step)"""
    from umap_b200.native import check, lib, ptr, stream
    mod = opt.mods[0]
    g = mod.graph
    mod.g.zero_()
    check(lib().mmu_edge_sample_range(ptr(g.row), ptr(g.col), ptr(g.val), 0, g.nnz, mod.batch_size, mod.n_batches, mod.seed,
                                      ptr(opt.state), ptr(mod.kept_rec), ptr(mod.kept_hdr), ptr(mod.batch_kept), stream()),
          "sample")
    opt._forces(mod, mod.kept_rec, mod.kept_hdr, None, mod.batch_kept)
    torch.cuda.synchronize()
    return mod.g.cpu().numpy().astype(np.float64), int(mod.kept_hdr[0].item())


@pytest.mark.parametrize("mode", ["fit", "transform"])
@pytest.mark.parametrize("dim,num_rep", [(2, 8), (4, 4), (8, 8), (16, 8), (16, 4), (64, 8), (128, 8)])
def test_staged_windowed_and_loop_kernels_agree(dim, num_rep, mode):
    """Device sample stream: the staged run-form kernel (records staged through shared memory, head gradient per
    run), the same kernel run in TAIL WINDOWS (one launch per window of tail rows, the Philox draws regenerated in
    every pass) and the plain loop kernel (option force_staged = 0) see the same kept edges and negatives and
    produce the same gradient, up to the order of the floating-point atomics."""
    from umap_b200 import native
    from umap_b200.layout import LayoutOptimizer
    rng = np.random.default_rng(dim + num_rep)
    n, k, bs = 3000, 12, 256
    y = (rng.standard_normal((n, dim)) * 0.3).astype(np.float32)
    ref = (rng.standard_normal((n + 100, dim)) * 0.3).astype(np.float32)
    cols = np.stack([np.sort(rng.choice(n, k, replace=False)) for _ in range(n)])
    graph = _coo(np.repeat(np.arange(n), k), cols.reshape(-1), rng.random(n * k).astype(np.float32), (n, n + (100 if mode == "transform" else 0)))

    def run(staged, window_rows):
        native.set_option("force_staged", staged)
        try:
            opt = LayoutOptimizer([torch.from_numpy(y)], [graph], 1.577, 0.8951, num_rep, 0.01, 1.0, bs, mode=mode,
                                  refs=[torch.from_numpy(ref)] if mode == "transform" else None, sample_stream="device", seed=5)
            opt.mods[0].window_rows = window_rows
            out = _device_epoch_gradient(opt)
            name = native.last_kernel("edge_forces")
        finally:
            native.set_option("force_staged", 1)
        return out, name

    (g_staged, n_staged), name = run(1, 0)
    assert "staged" in name and "one-pass" in name, name
    (g_win, n_win), name = run(1, 700)                    # 5 windows, the last one short
    assert "windowed" in name, name
    (g_loop, n_loop), name = run(0, 0)
    assert "edge_forces_kernel" in name, name
    assert n_staged == n_win == n_loop and n_staged > 0
    scale = np.abs(g_loop).max()
    assert np.abs(g_staged - g_loop).max() < 1e-4 * scale           # fast vs fast: only atomic order (both use fast math)
    assert np.abs(g_win - g_staged).max() < 2e-5 * scale


def test_modalities_draw_independent_streams():
    """two modalities with IDENTICAL graphs must not keep the same edges or draw the same negatives
    (the reference draws an independent rand / randint per modality, model.py:432,444)."""
    from umap_b200.layout import LayoutOptimizer
    from umap_b200.native import check, lib, ptr, stream
    rng = np.random.default_rng(8)
    n, k, dim = 4000, 10, 16
    y = (rng.standard_normal((n, dim)) * 0.3).astype(np.float32)
    cols = np.stack([np.sort(rng.choice(n, k, replace=False)) for _ in range(n)])
    graph = _coo(np.repeat(np.arange(n), k), cols.reshape(-1), np.full(n * k, 0.5, np.float32), (n, n))
    opt = LayoutOptimizer([torch.from_numpy(y), torch.from_numpy(y)], [graph, graph], 1.577, 0.8951, 8, 0.01, 1.0, 256,
                          mode="fit", sample_stream="device", seed=3)
    assert opt.mods[0].seed != opt.mods[1].seed
    kept = []
    for mod in opt.mods:
        g = mod.graph
        check(lib().mmu_edge_sample_range(ptr(g.row), ptr(g.col), ptr(g.val), 0, g.nnz, 256, mod.n_batches, mod.seed,
                                          ptr(opt.state), ptr(mod.kept_rec), ptr(mod.kept_hdr), ptr(mod.batch_kept), stream()), "s")
        kept.append(set(mod.kept_rec[: int(mod.kept_hdr[0].item()), 0].cpu().numpy().tolist()))
    both = len(kept[0] & kept[1])
    # independent Bernoulli(0.5) draws: |A & B| ~ nnz / 4, identical streams would give |A & B| = |A|
    assert abs(both - n * k / 4) < 6 * np.sqrt(n * k * 3 / 16), (both, len(kept[0]))


def test_infonce_device_stream_rotates_the_short_chunk():
    """with num % 1000 != 0 the rows of the short last chunk get a larger weight; the device stream must
    not pin that chunk to a fixed row set.  With e0 == e1 == constant rows every anchor's gradient norm is
    proportional to its weight: the heavy rows must differ between epochs."""
    from umap_b200.native import check, lib, ptr, stream
    num, dim = 1500, 16
    g_ = torch.Generator().manual_seed(1)
    e0 = torch.randn(num, dim, generator=g_).cuda()
    e1 = torch.randn(num, dim, generator=g_).cuda()
    state = torch.zeros(8, dtype=torch.int32, device="cuda")
    check(lib().mmu_opt_state_init(ptr(state), stream()), "init")
    heavy = []
    for _ in range(3):
        g0, g1 = torch.zeros_like(e0), torch.zeros_like(e1)
        loss = torch.zeros(1, device="cuda")
        # one direction, no negatives: the gradient of anchor row i is w_i * (projected positive direction)
        check(lib().mmu_infonce_range(ptr(e0), ptr(e1), num, 0, num, dim, None, None, 1, 1000, 1.0, 0.5, ptr(g0), ptr(g1),
                                      7, 0, ptr(state), ptr(loss), stream()), "nce")
        # which rows sit in the short chunk: replay the same call with chunk weights made visible through the loss
        # of single-anchor ranges is expensive; use the anchor-gradient norm relative to a chunk=num run instead
        f0, f1 = torch.zeros_like(e0), torch.zeros_like(e1)
        check(lib().mmu_infonce_range(ptr(e0), ptr(e1), num, 0, num, dim, None, None, 1, num, 1.0, 0.5, ptr(f0), ptr(f1),
                                      7, 0, ptr(state), ptr(loss), stream()), "nce")
        ratio = (g0.norm(dim=1) / f0.norm(dim=1).clamp(min=1e-30)).cpu().numpy()
        # weights: 1/(1000*2) for the full chunk, 1/(500*2) for the short one, against 1/1500 in the reference run
        rows = np.nonzero(ratio > 1.2)[0]
        assert 400 <= rows.size <= 500, rows.size            # the 500 short-chunk anchors (minus masked negatives)
        heavy.append(set(rows.tolist()))
        check(lib().mmu_opt_state_advance(ptr(state), 0.01, 0.9, 0.999, stream()), "advance")
    assert len(heavy[0] & heavy[1]) < 0.9 * len(heavy[0]) or len(heavy[1] & heavy[2]) < 0.9 * len(heavy[1])


def test_overlapped_sampling_equals_in_order_sampling(monkeypatch):
    """The side-stream sampler (epoch e+1 sampled while the forces of epoch e run, explicit epoch
    numbers, double-buffered kept lists) draws the same edges per epoch as the in-order path: after 6
    epochs the embeddings agree up to the order of the fp32 atomics, the kept counts exactly."""
    from umap_b200.layout import LayoutOptimizer
    rng = np.random.default_rng(2)
    n, k, dim = 20000, 12, 16
    y = (rng.standard_normal((n, dim)) * 0.05).astype(np.float32)
    cols = np.sort(rng.integers(0, n, (n, k)), axis=1)
    cols += np.arange(k)[None, :]                       # strictly increasing -> distinct columns per row
    cols %= n
    cols = np.sort(cols, axis=1)
    graph = _coo(np.repeat(np.arange(n), k), cols.reshape(-1), rng.random(n * k).astype(np.float32), (n, n))
    outs, kept = [], []
    for flag in ("1", "0"):
        monkeypatch.setenv("MMUMAP_OVERLAP_SAMPLE", flag)
        monkeypatch.setenv("MMUMAP_GRAPH", "0")
        opt = LayoutOptimizer([torch.from_numpy(y)], [graph], 1.577, 0.8951, 8, 0.01, 1.0, 256, mode="fit",
                              sample_stream="device", seed=11)
        outs.append(opt.run(6)[0].cpu().numpy())
        kept.append(opt.kept_last_epoch())
    assert kept[0] == kept[1]
    diff = np.abs(outs[0] - outs[1])
    assert diff.mean() < 1e-6 and diff.max() < 5e-3, (diff.mean(), diff.max())
