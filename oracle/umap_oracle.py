"""oracle/umap_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (numpy) restatement of the reference's UMAP hot path, stage by stage.  Every
function cites the lines of /root/reference/impl/model.py it follows.  The
restatement is pinned by tests/golden/*.npz, which were produced by importing and
running the reference itself in the build container (oracle/make_golden.py);
tests/test_oracle_golden.py checks every function here against those vectors.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this
module.  The product path (multimodal-umap_b200/) never does.

Parity pin status: the reference ships no tests or golden vectors of its own
(SURVEY.md section 4), so the pins are outputs of the reference run live here.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


# --------------------------------------------------------------------------- kNN
def _lib():
    """Load (building if needed) the C restatement in knn_oracle.c."""
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "_build", "liboracle.so")
        src = os.path.join(_HERE, "knn_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        lib = ctypes.CDLL(so)
        lib.oracle_knn_exact.restype = ctypes.c_int
        lib.oracle_knn_exact.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
            ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        lib.oracle_pair_dist.restype = ctypes.c_int
        lib.oracle_pair_dist.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int64, ctypes.c_void_p]
        _LIB = lib
    return _LIB


def knn_exact(query: np.ndarray, db: np.ndarray, k: int, exclude_self: bool, self_offset: int = 0,
              nthreads: int = 0):
    """Exhaustive kNN under the canonical fp32 distance (model.py:109,163) and the
    (distance, index) ascending selection rule (model.py:181-193, :88/:166 self-exclusion).
    Returns (idx int32 [Q,k], dist float32 [Q,k])."""
    query = np.ascontiguousarray(query, dtype=np.float32)
    db = np.ascontiguousarray(db, dtype=np.float32)
    Q, D = query.shape
    N = db.shape[0]
    idx = np.empty((Q, k), dtype=np.int32)
    dist = np.empty((Q, k), dtype=np.float32)
    rc = _lib().oracle_knn_exact(query.ctypes.data, Q, db.ctypes.data, N, D, k, int(exclude_self),
                                 int(self_offset), idx.ctypes.data, dist.ctypes.data, int(nthreads))
    if rc != 0:
        raise RuntimeError(f"oracle_knn_exact failed rc={rc}")
    return idx, dist


def pair_dist(query: np.ndarray, db: np.ndarray, qi: np.ndarray, dj: np.ndarray) -> np.ndarray:
    query = np.ascontiguousarray(query, dtype=np.float32)
    db = np.ascontiguousarray(db, dtype=np.float32)
    qi = np.ascontiguousarray(qi, dtype=np.int64)
    dj = np.ascontiguousarray(dj, dtype=np.int64)
    out = np.empty(qi.shape[0], dtype=np.float32)
    _lib().oracle_pair_dist(query.ctypes.data, db.ctypes.data, query.shape[1], qi.ctypes.data,
                            dj.ctypes.data, qi.shape[0], out.ctypes.data)
    return out


def knn_exact_numpy(query: np.ndarray, db: np.ndarray, k: int, exclude_self: bool, self_offset: int = 0):
    """Pure-numpy small-case version of knn_exact (same canonical order), used to
    cross-check the C code."""
    query = np.asarray(query, dtype=np.float32)
    db = np.asarray(db, dtype=np.float32)
    Q, D = query.shape
    N = db.shape[0]
    acc = np.zeros((Q, N), dtype=np.float32)
    for t in range(D):
        diff = (query[:, t:t + 1] - db[None, :, t]).astype(np.float32)
        # fmaf(diff, diff, acc): single rounding -> evaluate in float64 (exact product of
        # two fp32 and an fp32 addend fits 53 bits except in far-apart-exponent cases where
        # the double rounding coincides) and round once to fp32.
        acc = (diff.astype(np.float64) * diff.astype(np.float64) + acc.astype(np.float64)).astype(np.float32)
    dist = np.sqrt(acc).astype(np.float32)
    if exclude_self:
        q = np.arange(Q)
        j = q + self_offset
        ok = (j >= 0) & (j < N)
        dist[q[ok], j[ok]] = np.inf
    order = np.lexsort((np.broadcast_to(np.arange(N), (Q, N)), dist), axis=1)[:, :k]
    return order.astype(np.int32), np.take_along_axis(dist, order, axis=1)


# --------------------------------------------------------------------- sigma / rho
def sigmas_newton(dists: np.ndarray, num_iters: int = 20) -> np.ndarray:
    """model.py:33-61 restated with the closed-form derivative the reference obtains
    through autograd: d/dsigma sum_j exp(-(d_j-rho)/sigma) = sum_j p_j (d_j-rho)/sigma^2.
    fp32 throughout; start sigma=1; sigma <- clamp(sigma - val/(grad+1e-6), 1e-6)."""
    d = np.asarray(dists, dtype=np.float32)
    k = d.shape[1]
    rho = d.min(axis=1, keepdims=True)
    delta = (d - rho).astype(np.float32)
    target = np.float32(np.log2(np.float32(k)))
    sigma = np.ones(d.shape[0], dtype=np.float32)
    with np.errstate(over="ignore", under="ignore", divide="ignore", invalid="ignore"):
        for _ in range(num_iters):
            s = sigma[:, None]
            p = np.exp((-delta / s).astype(np.float32)).astype(np.float32)
            val = p.sum(axis=1, dtype=np.float32) - target
            # autograd: d(-delta/s)/ds = delta/(s*s); grad = sum p * delta/(s*s)
            grad = (p * (delta / (s * s)).astype(np.float32)).sum(axis=1, dtype=np.float32)
            sigma = np.maximum((sigma - val / (grad + np.float32(1e-6))).astype(np.float32),
                               np.float32(1e-6))
    return sigma


def sigmas_bisect(dists: np.ndarray, num_iters: int = 64, tol: float = 1e-5) -> np.ndarray:
    """Bisection solve of the same equation sum_j exp(-(d_j-rho)/sigma) = log2(k)
    (the equation at model.py:46-50).  This is the solver north_star asks the new engine
    to run; it equals the reference's Newton result on rows where Newton converged."""
    d = np.asarray(dists, dtype=np.float32)
    k = d.shape[1]
    rho = d.min(axis=1, keepdims=True)
    delta = (d - rho).astype(np.float32)
    target = np.float32(np.log2(np.float32(k)))
    n = d.shape[0]
    lo = np.zeros(n, dtype=np.float32)
    hi = np.full(n, np.inf, dtype=np.float32)
    mid = np.ones(n, dtype=np.float32)
    done = np.zeros(n, dtype=bool)
    with np.errstate(over="ignore", under="ignore", divide="ignore", invalid="ignore"):
        for _ in range(num_iters):
            s = np.exp((-delta / mid[:, None]).astype(np.float32)).sum(axis=1, dtype=np.float32)
            done |= np.abs(s - target) < np.float32(tol)
            gt = s > target
            new_hi = np.where(gt & ~done, mid, hi)
            new_lo = np.where(~gt & ~done, mid, lo)
            new_mid = np.where(np.isinf(new_hi), mid * np.float32(2.0),
                               ((new_lo + new_hi) * np.float32(0.5)).astype(np.float32))
            mid = np.where(done, mid, new_mid).astype(np.float32)
            lo, hi = new_lo, new_hi
    return np.maximum(mid, np.float32(1e-6))


def membership_weights(dists: np.ndarray, sigma: np.ndarray) -> np.ndarray:
    """model.py:199-201: rho = row minimum; w = exp(-(d-rho)/sigma)."""
    d = np.asarray(dists, dtype=np.float32)
    rho = d.min(axis=1, keepdims=True)
    with np.errstate(under="ignore"):
        return np.exp((-(d - rho) / sigma[:, None].astype(np.float32)).astype(np.float32)).astype(np.float32)


def invert_weights(dists: np.ndarray, a: float, b: float) -> np.ndarray:
    """model.py:206: w = 1/(1 + a d^(2b))."""
    d = np.asarray(dists, dtype=np.float32)
    return (np.float32(1.0) / (np.float32(1.0) + np.float32(a) * np.power(d, np.float32(2 * b)))).astype(np.float32)


def coalesce_rows(idx: np.ndarray, vals: np.ndarray):
    """model.py:208: sparse_coo_tensor(...).coalesce() orders each row's k entries by column."""
    order = np.argsort(idx, axis=1, kind="stable")
    return np.take_along_axis(idx, order, axis=1), np.take_along_axis(vals, order, axis=1)


# ------------------------------------------------------------------- fuzzy union
def fuzzy_union(rows: np.ndarray, cols: np.ndarray, vals: np.ndarray, n: int):
    """model.py:271: S = G + G^T - G*G^T, coalesced (sorted by row then col).
    Pattern = union of the patterns of G and G^T; value = fl(fl(a+b) - fl(a*b)) where both
    exist, else the single value (fp32, one rounding per operation as torch does)."""
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    vals = np.asarray(vals, dtype=np.float32)
    key_g = rows * n + cols
    key_t = cols * n + rows
    # G is coalesced: keys unique
    og = np.argsort(key_g, kind="stable")
    ot = np.argsort(key_t, kind="stable")
    kg, vg = key_g[og], vals[og]
    kt, vt = key_t[ot], vals[ot]
    allk = np.union1d(kg, kt)
    a = np.zeros(allk.shape[0], dtype=np.float32)
    b = np.zeros(allk.shape[0], dtype=np.float32)
    has_a = np.zeros(allk.shape[0], dtype=bool)
    has_b = np.zeros(allk.shape[0], dtype=bool)
    pa = np.searchsorted(allk, kg)
    pb = np.searchsorted(allk, kt)
    a[pa] = vg
    has_a[pa] = True
    b[pb] = vt
    has_b[pb] = True
    both = has_a & has_b
    out = np.where(has_a, a, b).astype(np.float32)
    s = (a[both] + b[both]).astype(np.float32)
    p = (a[both] * b[both]).astype(np.float32)
    out[both] = (s - p).astype(np.float32)
    return (allk // n).astype(np.int64), (allk % n).astype(np.int64), out


# ------------------------------------------------------------------ embed_query
def embed_query(rows: np.ndarray, cols: np.ndarray, vals: np.ndarray, q: int, ref: np.ndarray) -> np.ndarray:
    """model.py:236-252: row-normalised sparse (Q x N) @ dense (N x d)."""
    ref = np.asarray(ref, dtype=np.float32)
    sums = np.zeros(q, dtype=np.float32)
    np.add.at(sums, rows, vals.astype(np.float32))
    sums = np.maximum(sums, np.float32(1e-6))
    w = (vals / sums[rows]).astype(np.float32)
    out = np.zeros((q, ref.shape[1]), dtype=np.float32)
    np.add.at(out, rows, w[:, None] * ref[cols])
    return out


# ------------------------------------------------------------------ force terms
def _pow(x, p):
    return np.power(x, p)


def umap_attr_grad(y_i: np.ndarray, y_j: np.ndarray, a: float, b: float):
    """Closed form of model.py:312-322 for ONE batch: loss = mean_e log(1 + a s^b),
    s = clamp(|y_i-y_j|^2, 1e-6).  Returns (loss, dL/dy_i per edge); dL/dy_j = -dL/dy_i.
    Gradient is zero where the clamp is active (torch clamp backward)."""
    diff = y_i - y_j
    s_raw = (diff * diff).sum(axis=1)
    s = np.maximum(s_raw, 1e-6)
    n = y_i.shape[0]
    sb = _pow(s, b)
    loss = np.log(1.0 + a * sb).mean() if n else 0.0
    coef = 2.0 * a * b * sb / s / (1.0 + a * sb) / max(n, 1)
    coef = np.where(s_raw >= 1e-6, coef, 0.0)
    return loss, coef[:, None] * diff


def umap_rep_grad(y_i: np.ndarray, y_l: np.ndarray, a: float, b: float):
    """Closed form of model.py:324-334 for ONE batch:
    loss = mean -log(q/(1+q) + 1e-6), q = a s^b."""
    diff = y_i - y_l
    s_raw = (diff * diff).sum(axis=1)
    s = np.maximum(s_raw, 1e-6)
    n = y_i.shape[0]
    q = a * _pow(s, b)
    loss = (-np.log(q / (1.0 + q) + 1e-6)).mean() if n else 0.0
    coef = -2.0 * a * b * _pow(s, b) / s / ((q / (1.0 + q) + 1e-6) * (1.0 + q) ** 2) / max(n, 1)
    coef = np.where(s_raw >= 1e-6, coef, 0.0)
    return loss, coef[:, None] * diff


def inv_attr_grad(x_i: np.ndarray, y_j: np.ndarray, sigma_j: np.ndarray, a: float, b: float):
    """Closed form of model.py:336-348 (_inv_attr_loss) for ONE batch:
    loss = mean_e dist / (w sigma_j + 1e-6), s = clamp(|x_i - y_j|^2, 1e-6), dist = sqrt(s),
    w = 1/(1 + a s^b).  Returns (loss, dL/dx_i per edge); y (the data rows) are constants."""
    diff = x_i - y_j
    s_raw = (diff * diff).sum(axis=1)
    s = np.maximum(s_raw, 1e-6)
    dist = np.sqrt(s)
    n = x_i.shape[0]
    w = 1.0 / (1.0 + a * _pow(s, b))
    u = w * sigma_j + 1e-6
    loss = (dist / u).mean() if n else 0.0
    # d/ds [dist/u] = 1/(2 dist u) - dist * (dw/ds) sigma / u^2,   dw/ds = -a b s^(b-1) w^2
    dls = 1.0 / (2.0 * dist * u) + dist * sigma_j * a * b * _pow(s, b - 1.0) * w * w / (u * u)
    coef = np.where(s_raw >= 1e-6, 2.0 * dls, 0.0) / max(n, 1)
    return loss, coef[:, None] * diff


def inv_rep_grad(x_i: np.ndarray, y_l: np.ndarray, sigma_l: np.ndarray, rho_l: np.ndarray):
    """Closed form of model.py:350-362 (_inv_rep_loss) for ONE batch:
    loss = mean -log(1 - exp(-clamp(dist - rho_l, 1e-6)/(sigma_l + 1e-6)) + 1e-6)."""
    diff = x_i - y_l
    s_raw = (diff * diff).sum(axis=1)
    s = np.maximum(s_raw, 1e-6)
    dist = np.sqrt(s)
    n = x_i.shape[0]
    c_raw = dist - rho_l
    c = np.maximum(c_raw, 1e-6)
    e = np.exp(-c / (sigma_l + 1e-6))
    loss = (-np.log(1.0 - e + 1e-6)).mean() if n else 0.0
    # dL/de = 1/(1-e+1e-6); de/dc = -e/(sigma+1e-6); dc/ddist = [c_raw >= 1e-6]; ddist/ds = 1/(2 dist)
    dls = (1.0 / (1.0 - e + 1e-6)) * (-e / (sigma_l + 1e-6)) * (c_raw >= 1e-6) / (2.0 * dist)
    coef = np.where(s_raw >= 1e-6, 2.0 * dls, 0.0) / max(n, 1)
    return loss, coef[:, None] * diff


def infonce_grad(e0: np.ndarray, e1: np.ndarray, perm: np.ndarray, negs: np.ndarray,
                 temperature: float = 0.5, chunk: int = 1000):
    """Closed form of model.py:364-394 given the replayed draws: `perm` = the randperm
    (model.py:373) and `negs` [num, 9] = the per-chunk randint draws concatenated in
    anchor order (model.py:383).  Returns (loss, grad_e0, grad_e1)."""
    num = perm.shape[0]
    g0 = np.zeros_like(e0, dtype=np.float64)
    g1 = np.zeros_like(e1, dtype=np.float64)
    n_chunks = (num + chunk - 1) // chunk
    loss = 0.0
    for c in range(n_chunks):
        lo, hi = c * chunk, min((c + 1) * chunk, num)
        wgt = 1.0 / ((hi - lo) * n_chunks)
        for t in range(lo, hi):
            i = int(perm[t])
            av = e0[i].astype(np.float64)
            an = max(np.linalg.norm(av), 1e-12)
            u = av / an
            cand = [i] + [int(x) for x in negs[t]]
            valid = [True] + [int(x) != i for x in negs[t]]
            vs, ns, logits = [], [], []
            for m, (ci, ok) in enumerate(zip(cand, valid)):
                ev = e1[ci].astype(np.float64)
                en = max(np.linalg.norm(ev), 1e-12)
                v = ev / en
                vs.append(v)
                ns.append(en)
                logits.append(np.dot(u, v) / temperature if ok else -np.inf)
            logits = np.array(logits)
            mx = logits.max()
            ex = np.exp(logits - mx)
            pi = ex / ex.sum()
            loss += wgt * (-(logits[0] - mx - np.log(ex.sum())))
            cm = pi.copy()
            cm[0] -= 1.0
            acc = np.zeros_like(u)
            for m, (ci, ok) in enumerate(zip(cand, valid)):
                if not ok:
                    continue
                acc += cm[m] * vs[m]
                gv = cm[m] * u
                gv = gv - vs[m] * np.dot(vs[m], gv)
                g1[ci] += wgt * gv / (temperature * ns[m])
            ga = acc - u * np.dot(u, acc)
            g0[i] += wgt * ga / (temperature * an)
    return loss, g0, g1


def infonce_grad_vec(e0: np.ndarray, e1: np.ndarray, perm: np.ndarray, negs: np.ndarray,
                     temperature: float = 0.5, chunk: int = 1000):
    """Vectorised numpy form of infonce_grad (same closed form of model.py:364-394, fp64);
    used where the per-anchor Python loop is too slow (bench.py's CPU baseline).  Pinned to the
    loop version by tests/test_oracle_golden.py."""
    num = perm.shape[0]
    g0 = np.zeros(e0.shape, dtype=np.float64)
    g1 = np.zeros(e1.shape, dtype=np.float64)
    if num == 0:
        return 0.0, g0, g1
    n_chunks = (num + chunk - 1) // chunk
    t = np.arange(num)
    clen = np.minimum(chunk, num - (t // chunk) * chunk)
    wgt = 1.0 / (clen * n_chunks)                                    # [num]
    i = perm.astype(np.int64)
    cand = np.concatenate([i[:, None], negs.astype(np.int64)], axis=1)          # [num, M]
    valid = np.concatenate([np.ones((num, 1), bool), negs != i[:, None]], axis=1)
    a = e0[i].astype(np.float64)
    an = np.maximum(np.linalg.norm(a, axis=1), 1e-12)
    u = a / an[:, None]
    ev = e1[cand].astype(np.float64)                                  # [num, M, d]
    en = np.maximum(np.linalg.norm(ev, axis=2), 1e-12)
    v = ev / en[:, :, None]
    cs = np.einsum("nd,nmd->nm", u, v)
    logits = np.where(valid, cs / temperature, -np.inf)
    mx = logits.max(axis=1, keepdims=True)
    ex = np.exp(logits - mx)
    den = ex.sum(axis=1, keepdims=True)
    pi = ex / den
    loss = float((wgt * -(logits[:, 0] - mx[:, 0] - np.log(den[:, 0]))).sum())
    cm = pi.copy()
    cm[:, 0] -= 1.0
    cm = np.where(valid, cm, 0.0)
    acc = np.einsum("nm,nmd->nd", cm, v)
    ga = acc - u * (u * acc).sum(axis=1, keepdims=True)
    np.add.at(g0, i, (wgt / (temperature * an))[:, None] * ga)
    gv = cm[:, :, None] * u[:, None, :]
    gv = gv - v * (v * gv).sum(axis=2, keepdims=True)
    gv = gv * (wgt[:, None] / (temperature * en))[:, :, None]
    np.add.at(g1, cand.reshape(-1), gv.reshape(-1, e1.shape[1]))
    return loss, g0, g1


def adam_step(p, g, m, v, step: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor update (model.py:403,476), fp32."""
    p = p.astype(np.float32)
    g = g.astype(np.float32)
    m = (m + np.float32(1.0 - beta1) * (g - m)).astype(np.float32)
    v = (v * np.float32(beta2) + (np.float32(1.0 - beta2) * g) * g).astype(np.float32)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = np.float32(lr / bc1)
    bc2_sqrt = np.float32(bc2 ** 0.5)
    denom = (np.sqrt(v) / bc2_sqrt + np.float32(eps)).astype(np.float32)
    p = (p + (np.float32(-1.0) * step_size * m) / denom).astype(np.float32)
    return p, m, v


# ------------------------------------------------------------- _train restated
def train_oracle(embeds, graphs, epochs, num_rep, lr, alpha, batch_size, a, b, mode="fit",
                 refs=None, record=None, infonce=None):
    """Restatement of UMAPMixture._train (model.py:396-481) for modes "fit"/"transform",
    drawing from torch's global CPU generator in the reference's call order
    (SURVEY.md section 3.3).  `graphs` = list of (rows, cols, vals) coalesced COO arrays;
    `refs` = frozen reference tables in transform mode.  Gradients in fp64, state in fp32.
    Returns the list of final embeddings; per-epoch losses are appended to `record`."""
    import torch

    ys = [np.asarray(e, dtype=np.float32).copy() for e in embeds]
    ms = [np.zeros_like(y) for y in ys]
    vs = [np.zeros_like(y) for y in ys]
    for epoch in range(epochs):
        grads = [np.zeros(y.shape, dtype=np.float64) for y in ys]
        total = 0.0
        for i, y in enumerate(ys):
            rows, cols, vals = graphs[i]
            ref = refs[i] if mode == "transform" else None
            count = y.shape[0]
            nb = (count + batch_size - 1) // batch_size
            lsum = 0.0
            for bi, j in enumerate(range(0, count, batch_size)):
                end = min(j + batch_size, count)
                sel = (rows >= j) & (rows < end)
                br, bc, bv = rows[sel], cols[sel], vals[sel]
                keep = (torch.rand(bv.shape[0]).numpy() < bv)
                ii, jj = br[keep], bc[keep]
                tgt = ref if ref is not None else y
                la, ga = umap_attr_grad(y[ii].astype(np.float64), tgt[jj].astype(np.float64), a, b)
                np.add.at(grads[i], ii, ga / nb)
                if ref is None:
                    np.add.at(grads[i], jj, -ga / nb)
                npairs = ii.shape[0]
                rep_count = ref.shape[0] if ref is not None else count
                ll = torch.randint(0, rep_count, (npairs, num_rep)).flatten().numpy()
                ir = np.repeat(ii, num_rep)
                lr_, gr = umap_rep_grad(y[ir].astype(np.float64), tgt[ll].astype(np.float64), a, b)
                np.add.at(grads[i], ir, gr / nb)
                if ref is None:
                    np.add.at(grads[i], ll, -gr / nb)
                lsum += la + lr_
            total += lsum / nb
        if mode == "fit":
            n = len(ys)
            for i in range(n):
                for j in range(i + 1, n):
                    for (s, t) in ((i, j), (j, i)):
                        num = min(ys[s].shape[0], ys[t].shape[0])
                        perm = torch.randperm(num).numpy()
                        negs = []
                        for st in range(0, num, 1000):
                            en = min(st + 1000, num)
                            negs.append(torch.randint(0, num, (en - st, 9)).numpy())
                        negs = np.concatenate(negs, axis=0) if negs else np.zeros((0, 9), dtype=np.int64)
                        l, g0, g1 = (infonce or infonce_grad)(ys[s], ys[t], perm, negs)
                        # total loss carries alpha*(L_ij+L_ji): model.py:467-472
                        grads[s] += alpha * g0
                        grads[t] += alpha * g1
                        total += alpha * l
        if record is not None:
            record.append(total)
        for i in range(len(ys)):
            ys[i], ms[i], vs[i] = adam_step(ys[i], grads[i].astype(np.float32), ms[i], vs[i], epoch + 1, lr)
    return ys


# ---------------------------------------------------------------- spectral init
def laplacian_residual(rows, cols, vals, n, vecs):
    """Checks vectors against the operator of model.py:221-230:
    L = I - D^-1/2 S D^-1/2 + 1e-6 I.  Returns (rayleigh quotients, residual norms)."""
    import scipy.sparse as sp

    s = sp.coo_matrix((vals.astype(np.float64), (rows, cols)), shape=(n, n)).tocsr()
    deg = np.maximum(np.asarray(s.sum(axis=1)).ravel(), 1e-6)
    dm = sp.diags(deg ** -0.5)
    lap = sp.identity(n) * (1.0 + 1e-6) - dm @ s @ dm
    v = np.asarray(vecs, dtype=np.float64)
    lv = lap @ v
    lam = (v * lv).sum(axis=0) / (v * v).sum(axis=0)
    res = np.linalg.norm(lv - v * lam, axis=0)
    return lam, res
