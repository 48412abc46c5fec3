"""oracle/make_golden.py -- generates tests/golden/*.npz by RUNNING THE REFERENCE.

Run in the build container only (it reads /root/reference, which does not exist on the
GPU box):      python oracle/make_golden.py

The reference has no tests/golden vectors of its own, so these
fixtures are the pin: each file stores seeded inputs and the outputs the unmodified
reference functions (/root/reference/impl/model.py) produced for them on CPU with
torch 2.11.0.  Nothing from the reference's source is copied; it is imported.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
REF = os.environ.get("MMUMAP_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import umap_oracle as orc  # noqa: E402


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_model", os.path.join(REF, "impl", "model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def blobs(n, d, centers, seed, spread=5.0):
    g = torch.Generator().manual_seed(seed)
    c = torch.randn(centers, d, generator=g) * spread
    lab = torch.randint(0, centers, (n,), generator=g)
    x = c[lab] + torch.randn(n, d, generator=g)
    return x.float().numpy(), lab.numpy()


def bert_like(n, d, centers, seed):
    g = torch.Generator().manual_seed(seed)
    c = torch.randn(centers, d, generator=g) * 0.5
    lab = torch.randint(0, centers, (n,), generator=g)
    x = torch.tanh(c[lab] + 0.3 * torch.randn(n, d, generator=g))
    return x.float().numpy()


def coo_from_knn(idx, w):
    q, k = idx.shape
    ci, cw = orc.coalesce_rows(idx, w)
    rows = np.repeat(np.arange(q, dtype=np.int64), k)
    return rows, ci.reshape(-1).astype(np.int64), cw.reshape(-1).astype(np.float32)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    torch.set_num_threads(4)
    k = 15

    # ---- (1) a/b curve coefficients: model.py:587-618
    mix = ref.UMAPMixture(k_neighbors=k, out_dim=2, min_dist=0.1, num_encoders=2)
    np.savez(os.path.join(OUT, "ab.npz"), min_dist=0.1, a=mix.a, b=mix.b)
    a, b = mix.a, mix.b

    # ---- (2) sigma: model.py:33-61 and weights :199-201 on exact-kNN distances
    enc = ref.UMAPEncoder(k, 2)
    for name, x in (("blobs", blobs(600, 16, 6, 1)[0]), ("bert", bert_like(500, 48, 5, 2))):
        idx, dist = orc.knn_exact(x, x, k, True)
        d = torch.from_numpy(dist)
        mn = d.min(dim=1).values.unsqueeze(1).repeat(1, k)
        sig = enc.get_sigmas(d, mn)
        w = torch.exp(-(d - mn) / sig.unsqueeze(1))
        np.savez(os.path.join(OUT, f"sigma_{name}.npz"), x=x, idx=idx, dist=dist,
                 sigma=sig.numpy(), weights=w.numpy())

    # ---- (3) fuzzy union: model.py:271
    x, _ = blobs(400, 8, 4, 3)
    idx, dist = orc.knn_exact(x, x, k, True)
    sig = orc.sigmas_bisect(dist)
    w = orc.membership_weights(dist, sig)
    rows, cols, vals = coo_from_knn(idx, w)
    g = torch.sparse_coo_tensor(torch.from_numpy(np.stack([rows, cols])), torch.from_numpy(vals), (400, 400)).coalesce()
    u = (g + g.transpose(0, 1) - g * g.transpose(0, 1)).coalesce()
    np.savez(os.path.join(OUT, "union.npz"), rows=rows, cols=cols, vals=vals, n=400,
             out_rows=u.indices()[0].numpy(), out_cols=u.indices()[1].numpy(), out_vals=u.values().numpy())

    # ---- (4) embed_query: model.py:236-252
    gq = torch.Generator().manual_seed(4)
    refemb = torch.randn(400, 5, generator=gq)
    qrows = np.repeat(np.arange(50, dtype=np.int64), k)
    qcols = np.stack([np.sort(torch.randperm(400, generator=gq)[:k].numpy()) for _ in range(50)]).reshape(-1)
    qvals = torch.rand(50 * k, generator=gq).numpy().astype(np.float32)
    qg = torch.sparse_coo_tensor(torch.from_numpy(np.stack([qrows, qcols])), torch.from_numpy(qvals), (50, 400)).coalesce()
    eq = enc.embed_query(refemb, qg)
    np.savez(os.path.join(OUT, "embed_query.npz"), rows=qrows, cols=qcols, vals=qvals, ref=refemb.numpy(),
             out=eq.numpy())

    # ---- (5) loss terms + autograd gradients: model.py:312-334, :364-394
    gl = torch.Generator().manual_seed(5)
    y = (torch.randn(120, 4, generator=gl) * 0.7).requires_grad_(True)
    ii = torch.randint(0, 120, (300,), generator=gl)
    jj = torch.randint(0, 120, (300,), generator=gl)
    jj[:5] = ii[:5]                      # exercise the clamp (zero distance)
    la = mix._umap_attr_loss(y, ii, jj, a, b)
    ga = torch.autograd.grad(la, y)[0]
    lr_ = mix._umap_rep_loss(y, ii, jj, a, b)
    gr = torch.autograd.grad(lr_, y)[0]
    e0 = (torch.randn(2300, 6, generator=gl)).requires_grad_(True)
    e1 = (torch.randn(2500, 6, generator=gl)).requires_grad_(True)
    torch.manual_seed(55)
    li = mix._infonce_loss(e0, e1)
    gi0, gi1 = torch.autograd.grad(li, [e0, e1])
    np.savez(os.path.join(OUT, "losses.npz"), a=a, b=b, y=y.detach().numpy(), ii=ii.numpy(), jj=jj.numpy(),
             attr_loss=la.item(), attr_grad=ga.numpy(), rep_loss=lr_.item(), rep_grad=gr.numpy(),
             e0=e0.detach().numpy(), e1=e1.detach().numpy(), infonce_seed=55,
             infonce_loss=li.item(), infonce_g0=gi0.numpy(), infonce_g1=gi1.numpy())

    # ---- (6) _train, fit mode, two modalities (model.py:396-481)
    xa, _ = blobs(300, 12, 5, 6)
    xb, _ = blobs(260, 20, 5, 7)
    graphs_np, graphs_t, embeds0 = [], [], []
    for xm in (xa, xb):
        n = xm.shape[0]
        idx, dist = orc.knn_exact(xm, xm, k, True)
        w = orc.membership_weights(dist, orc.sigmas_bisect(dist))
        r, c, v = coo_from_knn(idx, w)
        ur, uc, uv = orc.fuzzy_union(r, c, v, n)
        graphs_np.append((ur, uc, uv))
        graphs_t.append(torch.sparse_coo_tensor(torch.from_numpy(np.stack([ur, uc])), torch.from_numpy(uv), (n, n)).coalesce())
        embeds0.append((torch.randn(n, 3, generator=gl) * 0.05).numpy().astype(np.float32))
    out = {"a": a, "b": b, "num_rep": 4, "lr": 0.01, "alpha": 1.0, "batch_size": 128, "seed": 66}
    for m in range(2):
        out[f"rows{m}"], out[f"cols{m}"], out[f"vals{m}"] = graphs_np[m]
        out[f"init{m}"] = embeds0[m]
    for ep in (1, 5, 20):
        torch.manual_seed(66)
        res = mix._train([torch.from_numpy(e) for e in embeds0], graphs_t, ep, 4, 0.01, 1.0, 128, mode="fit")
        for m in range(2):
            out[f"fit{ep}_{m}"] = res[m].detach().numpy()
    np.savez(os.path.join(OUT, "train_fit.npz"), **out)

    # ---- (7) _train, transform mode (model.py:399-401,415-416,443)
    nq = 90
    xq, _ = blobs(nq, 12, 5, 8)
    idx, dist = orc.knn_exact(xq, xa, k, False)
    w = orc.membership_weights(dist, orc.sigmas_bisect(dist))
    r, c, v = coo_from_knn(idx, w)
    tg = torch.sparse_coo_tensor(torch.from_numpy(np.stack([r, c])), torch.from_numpy(v), (nq, 300)).coalesce()
    fitted = torch.from_numpy(out["fit20_0"]).clone()
    mix.embeds = [fitted, torch.from_numpy(out["fit20_1"]).clone()]
    q0 = enc.embed_query(fitted, tg)
    tout = {"a": a, "b": b, "num_rep": 4, "lr": 0.01, "batch_size": 32, "seed": 77, "rows": r, "cols": c,
            "vals": v, "ref": fitted.numpy(), "init": q0.numpy()}
    for ep in (1, 10):
        torch.manual_seed(77)
        res = mix._train([q0], [tg], ep, 4, 0.01, 1.0, 32, mode="transform", data_indices=[0])
        tout[f"tr{ep}"] = res[0].detach().numpy()
    np.savez(os.path.join(OUT, "train_transform.npz"), **tout)

    # ---- (8) the reference's own NN-descent graph on a small set: recorded only to document
    #          that exact kNN is a superset-quality answer (recall of the reference <= 1).
    xs, _ = blobs(500, 16, 5, 9)
    torch.manual_seed(88)
    enc2 = ref.UMAPEncoder(k, 2)
    gref = enc2.fuzzy_knn_graph(torch.from_numpy(xs), "fit")
    eidx, _ = orc.knn_exact(xs, xs, k, True)
    rr, cc = gref.indices().numpy()
    hit = 0
    exact_sets = [set(row.tolist()) for row in eidx]
    for r_, c_ in zip(rr, cc):
        hit += int(c_ in exact_sets[r_])
    np.savez(os.path.join(OUT, "nndescent_recall.npz"), x=xs, ref_rows=rr, ref_cols=cc,
             recall=hit / float(eidx.size))
    print("reference NN-descent recall vs exact:", hit / float(eidx.size))
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
