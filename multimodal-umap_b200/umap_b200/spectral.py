"""Spectral initialisation (ref: /root/reference/impl/model.py:211-234, embed_all).

Operator: L = I - D^-1/2 S D^-1/2 + 1e-6 I ; result = eigenvectors of the out_dim+1 smallest
eigenvalues with the first dropped, unit norm, UNSCALED.

SURVEY.md section 8(f1) ranks a GPU-native block eigensolver as the first "next" component;
until then the solve is torch.lobpcg exactly as in the reference (same defaults), but on the
device and fed by the engine's CSR arrays.  `method="subspace"` is the engine's own
Chebyshev-filtered subspace iteration over the mmu_spmm_csr kernel.
"""
from __future__ import annotations

import os

import torch

from .graph import Graph, spmm


def normalized_adjacency(g: Graph) -> torch.Tensor:
    """values of D^-1/2 S D^-1/2 on the pattern of S (model.py:223-227)."""
    deg = torch.zeros(g.n_rows, dtype=torch.float32, device=g.val.device)
    deg.index_add_(0, g.row.long(), g.val)
    dinv = deg.clamp(min=1e-6).pow(-0.5)
    return g.val * dinv[g.row.long()] * dinv[g.col.long()]


def spectral_lobpcg(g: Graph, out_dim: int) -> torch.Tensor:
    n = g.n_rows
    dev = g.val.device
    aval = normalized_adjacency(g)
    idx = torch.stack([g.row.long(), g.col.long()])
    eye_idx = torch.arange(n, device=dev).repeat(2, 1)
    lap = torch.sparse_coo_tensor(torch.cat([eye_idx, idx], dim=1),
                                  torch.cat([torch.full((n,), 1.0 + 1e-6, device=dev), -aval]), (n, n)).coalesce()
    _, vecs = torch.lobpcg(lap, k=out_dim + 1, largest=False)           # model.py:232
    return vecs[:, 1:].contiguous()


def spectral_subspace(g: Graph, out_dim: int, iters: int = 40, degree: int = 8, seed: int = 0) -> torch.Tensor:
    """Chebyshev-filtered subspace iteration for the largest eigenpairs of A = D^-1/2 S D^-1/2
    (= smallest of L), block size out_dim+1 plus guard vectors, Rayleigh-Ritz each sweep."""
    n = g.n_rows
    dev = g.val.device
    aval = normalized_adjacency(g)
    m = out_dim + 1
    blk = min(n, m + max(4, m // 2))
    gen = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn((n, blk), generator=gen, device=dev, dtype=torch.float32)
    # spectrum of A lies in [-1, 1]; wanted end is near +1.  Filter damps [-1, cut].
    cut = 0.6
    e, c = (cut + 1.0) / 2.0, (cut - 1.0) / 2.0      # half-width, centre of the damped interval
    for _ in range(iters):
        x, _ = torch.linalg.qr(x)
        # three-term Chebyshev recurrence on (A - c I)/e
        t0 = x
        t1 = (spmm(g, x, aval) - c * x) / e
        for _ in range(2, degree + 1):
            t2 = 2.0 * (spmm(g, t1, aval) - c * t1) / e - t0
            t0, t1 = t1, t2
        x = t1
    x, _ = torch.linalg.qr(x)
    ax = spmm(g, x, aval)
    h = x.T @ ax
    h = 0.5 * (h + h.T)
    evals, evecs = torch.linalg.eigh(h)
    order = torch.argsort(evals, descending=True)[:m]
    v = x @ evecs[:, order]
    v = v / v.norm(dim=0, keepdim=True)
    return v[:, 1:].contiguous()


def spectral_init(g: Graph, out_dim: int, method: str | None = None) -> torch.Tensor:
    method = method or os.environ.get("MMUMAP_SPECTRAL", "lobpcg")
    if g.n_rows < 3 * (out_dim + 1):
        raise ValueError(f"spectral init needs at least {3 * (out_dim + 1)} points for out_dim={out_dim}")
    if method == "lobpcg":
        return spectral_lobpcg(g, out_dim)
    if method == "subspace":
        return spectral_subspace(g, out_dim)
    raise ValueError(f"unknown spectral method {method!r}")
