"""Pins oracle/ (the CPU restatement) against golden vectors produced by running the
reference itself (oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import umap_oracle as orc


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_knn_c_matches_numpy_restatement():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((150, 37)).astype(np.float32)
    x[10] = x[3]                     # duplicate points -> distance ties broken by index
    x[77] = x[3]
    i_c, d_c = orc.knn_exact(x, x, 15, True)
    i_n, d_n = orc.knn_exact_numpy(x, x, 15, True)
    assert np.array_equal(i_c, i_n)
    assert np.array_equal(d_c.view(np.uint32), d_n.view(np.uint32))
    q = rng.standard_normal((20, 37)).astype(np.float32)
    i_c, d_c = orc.knn_exact(q, x, 7, False)
    i_n, d_n = orc.knn_exact_numpy(q, x, 7, False)
    assert np.array_equal(i_c, i_n)
    assert np.array_equal(d_c.view(np.uint32), d_n.view(np.uint32))


def test_knn_agrees_with_torch_vector_norm_up_to_near_ties():
    """The canonical accumulation order differs from torch's vectorised vector_norm
    (model.py:109) only in the last ulps: neighbour sets agree except across fp32 near-ties."""
    rng = np.random.default_rng(1)
    x = rng.standard_normal((300, 64)).astype(np.float32)
    idx, dist = orc.knn_exact(x, x, 15, True)
    xt = torch.from_numpy(x)
    full = torch.linalg.vector_norm(xt[:, None, :] - xt[None, :, :], dim=2)
    full.fill_diagonal_(float("inf"))
    td, ti = torch.sort(full, dim=1, stable=True)
    ti = ti[:, :15].numpy()
    mism = idx != ti
    # any mismatch must be a swap of two candidates whose fp32 distances are within 4 ulp
    for r, c in zip(*np.nonzero(mism)):
        a = full[r, idx[r, c]].item()
        b = full[r, ti[r, c]].item()
        assert abs(a - b) <= 4 * np.spacing(np.float32(max(a, b)))
    assert mism.mean() < 0.01
    assert np.allclose(dist, td[:, :15].numpy(), rtol=1e-6)


@pytest.mark.parametrize("name", ["sigma_blobs.npz", "sigma_bert.npz"])
def test_sigma_newton_matches_reference(golden_dir, name):
    g = _load(golden_dir, name)
    sig = orc.sigmas_newton(g["dist"])
    ref = g["sigma"]
    # rows in the reference's divergent 2-cycle land on exactly reproducible values;
    # converged rows agree to fp32 rounding
    rel = np.abs(sig - ref) / np.maximum(np.abs(ref), 1e-12)
    assert np.quantile(rel, 0.99) < 1e-4, rel.max()
    assert (rel < 1e-3).mean() > 0.995
    w = orc.membership_weights(g["dist"], sig)
    assert np.abs(w - g["weights"]).mean() < 1e-5


@pytest.mark.parametrize("name", ["sigma_blobs.npz", "sigma_bert.npz"])
def test_sigma_bisect_equals_newton_where_newton_converged(golden_dir, name):
    g = _load(golden_dir, name)
    d = g["dist"]
    k = d.shape[1]
    ref = g["sigma"]
    rho = d.min(axis=1, keepdims=True)
    resid = np.abs(np.exp(-(d - rho) / ref[:, None]).sum(axis=1) - np.log2(k))
    conv = resid < 1e-3
    sig = orc.sigmas_bisect(d)
    if name == "sigma_blobs.npz":
        assert conv.sum() > 0          # the BERT-like set has no row where the reference's Newton converges
    assert np.allclose(sig[conv], ref[conv], rtol=2e-4)
    # and bisection solves the equation on every row that has a solution
    res_b = np.abs(np.exp(-(d - rho) / sig[:, None]).sum(axis=1) - np.log2(k))
    assert np.quantile(res_b, 0.999) < 1e-3


def test_union_matches_reference(golden_dir):
    g = _load(golden_dir, "union.npz")
    r, c, v = orc.fuzzy_union(g["rows"], g["cols"], g["vals"], int(g["n"]))
    assert np.array_equal(r, g["out_rows"])
    assert np.array_equal(c, g["out_cols"])
    # torch's coalesce sums the three terms {a, b, -ab} of a mutual edge in sort order, which is
    # not a defined association ((b-ab)+a on most entries, others elsewhere); the restatement
    # evaluates the literal expression fl(fl(a+b) - fl(ab)): indices bit-exact, values <= 2 ulp.
    ulp = np.abs(v.view(np.int32).astype(np.int64) - g["out_vals"].view(np.int32).astype(np.int64))
    assert ulp.max() <= 2
    n = int(g["n"])
    kin = g["rows"] * n + g["cols"]
    single = np.isin(g["out_rows"] * n + g["out_cols"], kin) ^ np.isin(g["out_cols"] * n + g["out_rows"], kin)
    assert single.sum() > 0
    assert np.array_equal(v[single].view(np.uint32), g["out_vals"][single].view(np.uint32))


def test_embed_query_matches_reference(golden_dir):
    g = _load(golden_dir, "embed_query.npz")
    out = orc.embed_query(g["rows"], g["cols"], g["vals"], 50, g["ref"])
    assert np.allclose(out, g["out"], rtol=1e-5, atol=1e-6)


def test_force_terms_match_reference_autograd(golden_dir):
    g = _load(golden_dir, "losses.npz")
    a, b = float(g["a"]), float(g["b"])
    y = g["y"].astype(np.float64)
    ii, jj = g["ii"], g["jj"]
    la, ga = orc.umap_attr_grad(y[ii], y[jj], a, b)
    grad = np.zeros_like(y)
    np.add.at(grad, ii, ga)
    np.add.at(grad, jj, -ga)
    assert abs(la - float(g["attr_loss"])) < 1e-5
    assert np.allclose(grad, g["attr_grad"], rtol=1e-4, atol=1e-6)
    lr_, gr = orc.umap_rep_grad(y[ii], y[jj], a, b)
    grad = np.zeros_like(y)
    np.add.at(grad, ii, gr)
    np.add.at(grad, jj, -gr)
    assert abs(lr_ - float(g["rep_loss"])) < 1e-4
    assert np.allclose(grad, g["rep_grad"], rtol=1e-4, atol=1e-5)


def test_invert_terms_match_reference_autograd(golden_dir):
    """model.py:336-362 (_inv_attr_loss / _inv_rep_loss) incl. both clamp branches."""
    g = _load(golden_dir, "invert_losses.npz")
    a, b = float(g["a"]), float(g["b"])
    x, data = g["x"].astype(np.float64), g["data"].astype(np.float64)
    ii, jj = g["ii"], g["jj"]
    sig, rho = g["sigma"].astype(np.float64), g["rho"].astype(np.float64)
    la, ga = orc.inv_attr_grad(x[ii], data[jj], sig[jj], a, b)
    grad = np.zeros_like(x)
    np.add.at(grad, ii, ga)
    assert abs(la - float(g["attr_loss"])) < 1e-4 * abs(float(g["attr_loss"]))
    assert np.allclose(grad, g["attr_grad"], rtol=2e-4, atol=1e-5 * np.abs(g["attr_grad"]).max())
    lr_, gr = orc.inv_rep_grad(x[ii], data[jj], sig[jj], rho[jj])
    grad = np.zeros_like(x)
    np.add.at(grad, ii, gr)
    assert abs(lr_ - float(g["rep_loss"])) < 1e-4
    assert np.allclose(grad, g["rep_grad"], rtol=2e-4, atol=1e-5 * np.abs(g["rep_grad"]).max())


def test_infonce_matches_reference_autograd(golden_dir):
    g = _load(golden_dir, "losses.npz")
    e0, e1 = g["e0"], g["e1"]
    num = min(e0.shape[0], e1.shape[0])
    torch.manual_seed(int(g["infonce_seed"]))
    perm = torch.randperm(num).numpy()
    negs = np.concatenate([torch.randint(0, num, (min(s + 1000, num) - s, 9)).numpy()
                           for s in range(0, num, 1000)])
    loss, g0, g1 = orc.infonce_grad(e0, e1, perm, negs)
    assert abs(loss - float(g["infonce_loss"])) < 1e-5
    assert np.allclose(g0, g["infonce_g0"], rtol=1e-4, atol=1e-8)
    assert np.allclose(g1, g["infonce_g1"], rtol=1e-4, atol=1e-8)


def test_infonce_vectorised_equals_loop_version():
    rng = np.random.default_rng(4)
    e0 = rng.standard_normal((2300, 6)).astype(np.float32)
    e1 = rng.standard_normal((2500, 6)).astype(np.float32)
    num = 2300
    perm = rng.permutation(num)
    negs = rng.integers(0, num, (num, 9))
    negs[::7, 3] = perm[::7]                      # masked negatives (model.py:386)
    l0, a0, b0 = orc.infonce_grad(e0, e1, perm, negs)
    l1, a1, b1 = orc.infonce_grad_vec(e0, e1, perm, negs)
    assert abs(l0 - l1) < 1e-12
    assert np.allclose(a0, a1, rtol=1e-10, atol=1e-16) and np.allclose(b0, b1, rtol=1e-10, atol=1e-16)


def _graphs(g, n):
    return [(g[f"rows{m}"], g[f"cols{m}"], g[f"vals{m}"]) for m in range(n)]


@pytest.mark.parametrize("epochs,tol", [(1, 2e-6), (5, 2e-5)])
def test_train_fit_matches_reference(golden_dir, epochs, tol):
    """Short-horizon elementwise parity of the restated _train (model.py:396-481); the
    tolerance floor is the reference's own rerun spread."""
    g = _load(golden_dir, "train_fit.npz")
    torch.manual_seed(int(g["seed"]))
    out = orc.train_oracle([g["init0"], g["init1"]], _graphs(g, 2), epochs, int(g["num_rep"]),
                           float(g["lr"]), float(g["alpha"]), int(g["batch_size"]), float(g["a"]),
                           float(g["b"]), mode="fit")
    for m in range(2):
        assert np.abs(out[m] - g[f"fit{epochs}_{m}"]).max() < tol


def test_train_transform_matches_reference(golden_dir):
    g = _load(golden_dir, "train_transform.npz")
    torch.manual_seed(int(g["seed"]))
    out = orc.train_oracle([g["init"]], [(g["rows"], g["cols"], g["vals"])], 1, int(g["num_rep"]),
                           float(g["lr"]), 1.0, int(g["batch_size"]), float(g["a"]), float(g["b"]),
                           mode="transform", refs=[g["ref"]])
    assert np.abs(out[0] - g["tr1"]).max() < 2e-6


def test_ab_coefficients_recorded(golden_dir):
    g = _load(golden_dir, "ab.npz")
    assert abs(float(g["a"]) - 1.577) < 1e-2 and abs(float(g["b"]) - 0.8951) < 1e-2
