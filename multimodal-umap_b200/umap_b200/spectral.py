"""Spectral initialisation (ref: /root/reference/impl/model.py:211-234, embed_all).

Operator: L = I - D^-1/2 S D^-1/2 + 1e-6 I ; result = eigenvectors of the out_dim+1 smallest
eigenvalues with the first dropped, unit norm, UNSCALED.

SURVEY.md section 8(f1) ranks a GPU-native block eigensolver as the first "next" component;
`spectral_chebfsi` (default) is the engine's own Chebyshev-filtered subspace iteration over the
mmu_spmm_csr_axpby kernel; `method="lobpcg"` runs torch.lobpcg exactly as the reference does (same
defaults), on the device and fed by the engine's CSR arrays.
"""
from __future__ import annotations

import os

import torch

from . import dist as D
from .graph import Graph, spmm, spmm_axpby
from .native import check, lib, ptr, stream


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


def _gram(x: torch.Tensor, y: torch.Tensor, chunk: int = 256) -> torch.Tensor:
    """x^T y for tall-skinny [n_pad, b] blocks (n_pad a multiple of `chunk`) as one batched GEMM
    over row chunks + a reduction: cuBLAS' large-k sgemm is ~20x slower on this shape."""
    c = x.shape[0] // chunk
    return torch.bmm(x.view(c, chunk, x.shape[1]).transpose(1, 2), y.view(c, chunk, y.shape[1])).sum(0)


EIGH_DEVICE_MAX = 64


def _eigh_small(a: torch.Tensor):
    """Symmetric eigendecomposition of a b x b (b ~ 32) matrix, eigenvalues ascending: the engine's
    one-CTA Jacobi kernel (mmu_eigh_small) -- no host round trip, no stream synchronisation; cuSOLVER's
    syevd launch sequence is ~4x slower than even a LAPACK round trip at this size.  Blocks wider than
    the kernel's limit (out_dim > 55) go through LAPACK on the host."""
    n = a.shape[0]
    if n > EIGH_DEVICE_MAX or os.environ.get("MMUMAP_EIGH") == "host":
        lam, v = torch.linalg.eigh(a.cpu())
        return lam.to(a.device), v.to(a.device)
    a = a.contiguous()
    lam = torch.empty(n, dtype=torch.float32, device=a.device)
    v = torch.empty((n, n), dtype=torch.float32, device=a.device)
    check(lib().mmu_eigh_small(ptr(a), n, ptr(lam), ptr(v), stream()), "mmu_eigh_small")
    return lam, v


def _orthonormalise(x: torch.Tensor) -> torch.Tensor:
    """SVQB: x (x^T x)^-1/2 through a small symmetric eigendecomposition (no host sync, tolerant of
    nearly dependent columns, which the Chebyshev filter produces by design)."""
    gm = _gram(x, x)
    gm = 0.5 * (gm + gm.T)
    lam, v = _eigh_small(gm)
    lam = lam.clamp(min=lam.max() * 1e-10)
    return x @ (v * lam.rsqrt())


BLOCK_WIDTHS = (8, 16, 32)          # block widths the device-resident block operations are instantiated for


def block_width(out_dim: int) -> int:
    """Columns of the iteration block: the out_dim + 1 wanted vectors plus guard vectors (the convergence rate of a
    filtered subspace iteration is set by the gap to the first eigenvalue OUTSIDE the block)."""
    m = out_dim + 1
    return 8 if m <= 5 else 16 if m <= 12 else 32 if m <= 24 else -(-(m + 8) // 4) * 4


def spectral_chebfsi_device(g: Graph, out_dim: int, tol: float = 3e-4, max_iter: int = 30, degree: int = 10,
                            shard: bool = False) -> torch.Tensor:
    """Chebyshev-filtered subspace iteration with EVERY step on the engine's own kernels (csrc/block_eig.cu) and every
    scalar of the iteration (Ritz values, residual, convergence flag, filter edge, Chebyshev coefficients) in a
    device-resident control block: the host enqueues whole outer iterations and reads the flag once per chunk of
    iterations instead of once per iteration.  Same mathematics and output contract as `spectral_chebfsi` below.

    One outer iteration = 1 SpMM (A x) + Gram (x^T A x) + 32x32 Jacobi + fused rotation of x and A x with the column
    residuals + the Ritz bookkeeping kernel + `degree` fused three-term SpMM steps + two SVQB orthonormalisations
    (Gram scaled to unit diagonal -> Jacobi -> T = D^-1 V Lambda^-1/2 -> rotation): ~27 launches of this library."""
    n, dev = g.n_rows, g.val.device
    m = out_dim + 1
    b = block_width(out_dim)
    L = lib()
    st = stream()
    aval = normalized_adjacency(g)
    deg = torch.zeros(n, dtype=torch.float32, device=dev)
    deg.index_add_(0, g.row.long(), g.val)
    # Multi-GPU, large graphs: the operator applications (all but a few per cent of the solve) are sharded by rows -- rank r
    # computes rows row_block(r) of every product and the blocks are all-gathered in place (NCCL over NVLink); everything
    # else is replicated on the identical full blocks, so the ranks' results are bit-identical and nothing is reduced.
    world, rank = (D.world(), D.rank()) if shard else (1, 0)
    per = D.block_size(n, world) if world > 1 else n
    n_alloc = per * world
    r_lo, r_hi = (min(n, rank * per), min(n, rank * per + per)) if world > 1 else (0, n)
    bufs = [torch.zeros((n_alloc, b), dtype=torch.float32, device=dev) for _ in range(4)]
    x, ax, y0, y1 = bufs
    x[:n].copy_(torch.randn((n, b), dtype=torch.float32, device=dev))
    if world > 1:
        D.broadcast(x, 0)                                   # one start block for all ranks
    x[:n, 0] = deg.clamp(min=1e-6).sqrt()                   # the known top eigenvector of A
    ctl = torch.empty(L.mmu_block_ctl_words(), dtype=torch.float32, device=dev)
    flag = ctl.view(torch.int32)
    check(L.mmu_block_ctl_init(ptr(ctl), st), "mmu_block_ctl_init")
    ws = torch.empty(L.mmu_block_gram_workspace_bytes(b), dtype=torch.uint8, device=dev)
    gm = torch.empty((b, b), dtype=torch.float32, device=dev)
    tm = torch.empty((b, b), dtype=torch.float32, device=dev)
    vm = torch.empty((b, b), dtype=torch.float32, device=dev)
    lam = torch.empty(b, dtype=torch.float32, device=dev)
    dinv = torch.empty(b, dtype=torch.float32, device=dev)
    rp, col = ptr(g.rowptr), ptr(g.col)

    def orthonormalise(src, dst, scaled):
        """scaled: dst = src (src^T src)^-1/2 (SVQB) with the Gram matrix first brought to unit diagonal, which is the
        column equalisation the filtered block needs (p(theta_j) spans orders of magnitude).  Not scaled: the second,
        polishing pass over an already nearly orthonormal block, done as Cholesky-QR."""
        check(L.mmu_block_gram(ptr(src), ptr(src), n, b, 2 if scaled else 1, ptr(ws), ptr(gm), ptr(dinv), ptr(ctl), st), "gram")
        if scaled:
            check(L.mmu_eigh_small_flag(ptr(gm), b, ptr(lam), ptr(vm), ptr(flag), st), "eigh")
            check(L.mmu_block_svqb(ptr(lam), ptr(vm), ptr(dinv), b, ptr(tm), ptr(ctl), st), "svqb")
        else:
            # the block is already nearly orthonormal (an SVQB pass ran just before): Cholesky-QR, a few microseconds
            check(L.mmu_block_cholqr(ptr(gm), b, ptr(tm), ptr(ctl), st), "cholqr")
        check(L.mmu_block_rotate(ptr(src), ptr(dst), None, None, n, b, ptr(tm), 0, None, ptr(ctl), st), "rotate")

    orthonormalise(x, x, True)

    def spmm(src, slot, z, dst, what):
        if world == 1:
            check(L.mmu_block_spmm(rp, col, ptr(aval), n, ptr(src), b, ptr(ctl), slot, ptr(z), ptr(dst), st), what)
            return
        check(L.mmu_block_spmm_rows(rp, col, ptr(aval), r_lo, r_hi, ptr(src), b, ptr(ctl), slot, ptr(z), ptr(dst), st), what)
        import torch.distributed as tdist
        tdist.all_gather_into_tensor(dst, dst[rank * per: rank * per + per])          # in place: my block is already there

    def outer():
        nonlocal x, ax, y0, y1
        spmm(x, 0, None, ax, "spmm")
        check(L.mmu_block_gram(ptr(x), ptr(ax), n, b, 1, ptr(ws), ptr(gm), None, ptr(ctl), st), "gram")
        check(L.mmu_eigh_small_flag(ptr(gm), b, ptr(lam), ptr(vm), ptr(flag), st), "eigh")
        # x <- x V, ax <- ax V (Ritz vectors, descending), column residuals |A x_j - theta_j x_j|^2 into the control block
        check(L.mmu_block_rotate(ptr(x), ptr(x), ptr(ax), ptr(ax), n, b, ptr(vm), 1, ptr(lam), ptr(ctl), st), "rotate")
        check(L.mmu_block_ritz(ptr(lam), b, m, tol, max_iter, ptr(ctl), st), "ritz")
        # Chebyshev filter of degree `degree` on [-1, cut]: T_1 from x, T_2 with z = x, then z aliases the output
        spmm(x, 1, None, y1, "cheb1")
        spmm(y1, 2, x, y0, "cheb2")
        cur, prev = y0, y1
        for _ in range(3, degree + 1):
            spmm(cur, 2, prev, prev, "chebk")
            cur, prev = prev, cur
        # equalise the column lengths (unit-diagonal Gram) and orthonormalise twice; the result becomes the new block.
        # After convergence every kernel above and below returns at once, so x keeps the converged Ritz vectors.
        orthonormalise(cur, prev, True)
        orthonormalise(prev, x, False)

    chunk = 4
    iters = 0
    while True:
        for _ in range(chunk):
            outer()
        done, iters = flag[:2].tolist()                      # ONE synchronisation per chunk of outer iterations
        if done or iters >= max_iter:
            break
        chunk = 2
    if os.environ.get("MMUMAP_SPECTRAL_DEBUG") == "1":
        print(f"  chebfsi(device) n={n} b={b}: {iters} Rayleigh-Ritz steps, residual {float(ctl[2]):.3e}, cut {float(ctl[3]):.4f}")
    vecs = x[:n, 1:m]
    return (vecs / vecs.norm(dim=0, keepdim=True)).contiguous()


def spectral_chebfsi(g: Graph, out_dim: int, tol: float = 3e-4, max_iter: int = 30, degree: int = 10) -> torch.Tensor:
    """Chebyshev-filtered subspace iteration (Zhou-Saad) for the out_dim+1 largest eigenpairs of
    A = D^-1/2 S D^-1/2, i.e. the smallest of the reference's L = I - A + 1e-6 I (model.py:221-230),
    on the engine's own kernels: one fused SpMM launch per Chebyshev step (mmu_spmm_csr_axpby),
    batched-GEMM Gram matrices, 32x32 dense eigenproblems.  Stops when every wanted Ritz pair has
    |A x - theta x| < tol (torch.lobpcg's default tolerance is sqrt(eps_fp32) = 3.5e-4).  Same
    output contract as embed_all: unit-norm columns, first (trivial) vector dropped, unscaled."""
    debug = os.environ.get("MMUMAP_SPECTRAL_DEBUG") == "1"
    if debug:
        import time
        torch.cuda.synchronize()
        t_start = time.perf_counter()
    n, dev = g.n_rows, g.val.device
    m = out_dim + 1
    b = 32 if m <= 24 else -(-(m + 8) // 4) * 4
    n_pad = -(-n // 256) * 256
    aval = normalized_adjacency(g)
    x = torch.zeros((n_pad, b), dtype=torch.float32, device=dev)
    x[:n] = torch.randn((n, b), dtype=torch.float32, device=dev)
    deg = torch.zeros(n, dtype=torch.float32, device=dev)
    deg.index_add_(0, g.row.long(), g.val)
    x[:n, 0] = deg.clamp(min=1e-6).sqrt()                    # the known top eigenvector of A
    x = _orthonormalise(x)
    ax = torch.zeros_like(x)
    y0 = torch.zeros_like(x)
    y1 = torch.zeros_like(x)
    cut = 0.0
    for it in range(max_iter):
        spmm_axpby(g, aval, x, 1.0, 0.0, None, 0.0, out=ax)
        h = _gram(x, ax)
        theta, v = _eigh_small(0.5 * (h + h.T))
        theta, v = theta.flip(0), v.flip(1)                        # descending
        x = x @ v
        ax = ax @ v
        res = (ax[:, :m] - x[:, :m] * theta[:m]).norm(dim=0).max()
        res, lo, theta_m = torch.stack([res, theta[b - 1], theta[m - 1]]).tolist()   # one sync per iteration
        if os.environ.get("MMUMAP_SPECTRAL_DEBUG") == "1":
            print(f"  chebfsi it={it} res={res:.3e} theta[0]={float(theta[0]):.6f} theta[m-1]={theta_m:.6f} "
                  f"theta[b-1]={lo:.6f} cut={cut:.4f}")
        if res < tol or it == max_iter - 1:
            break
        # damp [-1, cut].  The classical choice is the smallest Ritz value of the block (-> lambda_b from
        # below); when a cluster of (near-)equal eigenvalues is wider than the block -- UMAP graphs of
        # well separated clusters have one eigenvalue ~1 per cluster -- that value runs into the wanted
        # ones and the filter stops filtering, so the edge is kept a fixed distance below them
        cut = max(min(lo, theta_m - 0.05), -0.5)
        e, c = (cut + 1.0) / 2.0, (cut - 1.0) / 2.0
        spmm_axpby(g, aval, x, 1.0 / e, -c / e, None, 0.0, out=y1)           # T_1
        y0.copy_(x)
        for _ in range(2, degree + 1):
            spmm_axpby(g, aval, y1, 2.0 / e, -2.0 * c / e, y0, -1.0, out=y0)  # T_k -> overwrites T_{k-2}
            y0, y1 = y1, y0
        # the block holds Ritz vectors, so the filtered columns p(A) x_j stay nearly orthogonal and differ
        # mainly in length (p(theta_j) spans orders of magnitude): equalise the lengths first, otherwise
        # the fp32 Gram matrix loses the weakly amplified (wanted, slowly converging) directions
        y1 = y1 / y1.norm(dim=0, keepdim=True).clamp(min=1e-30)
        x = _orthonormalise(_orthonormalise(y1))
    vecs = x[:n, 1:m]
    out = (vecs / vecs.norm(dim=0, keepdim=True)).contiguous()
    if debug:
        torch.cuda.synchronize()
        print(f"  chebfsi n={n}: {it + 1} iterations, {(time.perf_counter() - t_start) * 1e3:.1f} ms wall")
    return out


# Multi-GPU: the operator applications are sharded over the ranks (and all-gathered) only when the iteration block no
# longer fits the L2, i.e. when a single GPU's SpMM gathers go to DRAM: measured at 8 GPUs, 10M rows x 8 columns (320 MB)
# 441 -> 129 ms, but 1M rows x 32 columns (128 MB, L2 resident: 0.6 ms per SpMM against ~1 ms per all-gather) 52 -> 95 ms.
SHARD_MIN_BLOCK_BYTES = 256 << 20


def shardable(g: Graph, out_dim: int, method: str | None = None) -> bool:
    """True when spectral_init(..., shard=True) would run as a collective of all ranks (every rank must then call it)."""
    method = method or os.environ.get("MMUMAP_SPECTRAL", "chebfsi")
    return (D.world() > 1 and method == "chebfsi" and block_width(out_dim) in BLOCK_WIDTHS
            and g.n_rows * block_width(out_dim) * 4 >= SHARD_MIN_BLOCK_BYTES
            and os.environ.get("MMUMAP_SPECTRAL_SHARD", "1") == "1")


def spectral_init(g: Graph, out_dim: int, method: str | None = None, shard: bool = False) -> torch.Tensor:
    method = method or os.environ.get("MMUMAP_SPECTRAL", "chebfsi")
    if g.n_rows < 3 * (out_dim + 1):
        raise ValueError(f"spectral init needs at least {3 * (out_dim + 1)} points for out_dim={out_dim}")
    if method == "lobpcg":
        return spectral_lobpcg(g, out_dim)
    if method == "chebfsi":
        # device-resident form for the instantiated block widths (out_dim <= 23); the torch-assisted form otherwise
        if block_width(out_dim) in BLOCK_WIDTHS:
            return spectral_chebfsi_device(g, out_dim, shard=shard and shardable(g, out_dim, method))
        return spectral_chebfsi(g, out_dim)
    if method == "chebfsi_torch":
        return spectral_chebfsi(g, out_dim)
    raise ValueError(f"unknown spectral method {method!r}")
