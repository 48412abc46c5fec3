"""GPU tests of the spectral initialisation (role of UMAPEncoder.embed_all, model.py:211-234): the
engine's Chebyshev-filtered subspace iteration and the torch.lobpcg path must both return unit-norm
vectors that are eigenvectors of the reference's operator L = I - D^-1/2 S D^-1/2 + 1e-6 I for its
smallest non-trivial eigenvalues (checked against a dense eigensolve of the same operator; the
reference's own output is a random-start LOBPCG result, so the comparison is on eigenvalues,
residuals and the spanned subspace, not on entries)."""
import numpy as np
import pytest
import torch

from oracle import umap_oracle as orc

pytestmark = pytest.mark.gpu


def _graph(n, d, k, seed, centers):
    from umap_b200 import graph as G
    rng = np.random.default_rng(seed)
    c = rng.standard_normal((centers, d)) * 3.0
    x = (c[rng.integers(0, centers, n)] + rng.standard_normal((n, d))).astype(np.float32)
    xt = torch.from_numpy(x).cuda()
    idx, dist = G.knn_graph(xt, xt, k, True)
    col, w, _, _ = G.smooth_knn(idx, dist, "bisect")
    return G.fuzzy_union(col, w)


@pytest.mark.parametrize("method", ["chebfsi", "chebfsi_torch", "lobpcg"])
@pytest.mark.parametrize("n,out_dim,centers", [(1500, 2, 3), (2500, 8, 5)])
def test_spectral_init_solves_the_reference_operator(method, n, out_dim, centers):
    from umap_b200.spectral import spectral_init
    torch.manual_seed(0)
    g = _graph(n, 24, 15, n, centers)
    v = spectral_init(g, out_dim, method=method).cpu().numpy().astype(np.float64)
    assert v.shape == (n, out_dim)
    assert np.allclose(np.linalg.norm(v, axis=0), 1.0, atol=1e-4)                   # unit norm, unscaled (model.py:232-234)
    rows, cols, vals = (t.cpu().numpy() for t in (g.row, g.col, g.val))
    lam, res = orc.laplacian_residual(rows.astype(np.int64), cols.astype(np.int64), vals, n, v)
    assert res.max() < 2e-3, res
    # dense eigensolve of the same operator
    import scipy.sparse as sp
    s = sp.coo_matrix((vals.astype(np.float64), (rows, cols)), shape=(n, n)).toarray()
    dm = np.maximum(s.sum(1), 1e-6) ** -0.5
    lap = np.eye(n) * (1.0 + 1e-6) - dm[:, None] * s * dm[None, :]
    ew, ev = np.linalg.eigh(lap)
    # Rayleigh quotients match the smallest non-trivial eigenvalues (first one dropped, model.py:234)
    assert np.allclose(np.sort(lam), ew[1:out_dim + 1], atol=2e-3), (np.sort(lam), ew[1:out_dim + 1])
    # and the vectors lie in the span of the eigenvectors below the next spectral gap
    upto = int(np.searchsorted(ew, ew[out_dim] + 5e-3, side="right"))
    proj = ev[:, :upto].T @ v
    assert np.all(np.linalg.norm(proj, axis=0) > 0.98)


@pytest.mark.parametrize("n", [1, 2, 7, 17, 32, 48, 64])
def test_eigh_small_matches_lapack(n):
    """mmu_eigh_small (one-CTA Jacobi) against float64 LAPACK: eigenvalues, orthonormality,
    reconstruction; a Gram matrix with a wide spectrum (what SVQB hands it) and an indefinite
    Rayleigh-Ritz matrix."""
    from umap_b200.spectral import _eigh_small
    rng = np.random.default_rng(n)
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    mats = [q @ np.diag(np.logspace(0, -7, n)) @ q.T, q @ np.diag(np.linspace(-0.8, 1.0, n)) @ q.T,
            np.eye(n), np.zeros((n, n))]
    for a in mats:
        a = 0.5 * (a + a.T)
        lam, v = _eigh_small(torch.from_numpy(a.astype(np.float32)).cuda())
        torch.cuda.synchronize()
        lam, v = lam.cpu().numpy().astype(np.float64), v.cpu().numpy().astype(np.float64)
        ref = np.linalg.eigvalsh(a)
        scale = max(np.abs(ref).max(), 1e-30)
        assert np.all(np.diff(lam) >= 0)                                            # ascending
        # fp32 cyclic Jacobi: ~6 sweeps x (n-1) steps of rotations, a few 1e-6 |A| of accumulated rounding
        assert np.abs(lam - ref).max() <= 2e-5 * scale, (n, np.abs(lam - ref).max() / scale)
        assert np.abs(v.T @ v - np.eye(n)).max() < 5e-5
        assert np.abs(v @ np.diag(lam) @ v.T - a).max() <= 5e-5 * scale
