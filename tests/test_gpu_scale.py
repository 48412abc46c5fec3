"""Parity at BASELINE.json sizes: the tensor-core kNN path on the real C2 generators at
full size, a C3-shaped 200k x 768 slice and a C4-shaped 1M x 128, k=30 slice, compared -- indices AND fp32
distance bit patterns -- with the exhaustive CUDA-core kernel (mmu_knn_exact_f32, itself held bit-exact to
oracle/knn_oracle.c at small sizes by tests/test_gpu_graph.py) and, on a few hundred random rows, with the C
oracle directly (restatement of /root/reference/impl/model.py:109,163,181-193).  The certification of a row
rests on a model of tcgen05's fp32 accumulation error (knn_tc.cu: c_rel); these are the sizes where a wrong
model would show.  Also asserts which certification level each problem needed.

And C1 as such (BASELINE.json configs[0]: 2,000 x 64 blobs, k=15, 2-D): fitted end to end and held to the
trustworthiness@15 the reference pipeline reaches on the same data."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from oracle import umap_oracle as orc

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _bench():
    return importlib.import_module("bench")


def _tc(q, db, k, excl):
    from umap_b200 import knn_tc
    i, d = knn_tc.knn_tc(q, db, k, excl)
    torch.cuda.synchronize()
    return i, d, dict(knn_tc.last_stats)


def _check_rows_vs_exact(x, ti, td, k, rows):
    """tensor-core result on `rows` vs the exhaustive kernel on the same rows (fit mode, self excluded)"""
    from umap_b200 import graph as G
    ei = torch.full_like(ti, -7)
    ed = torch.full_like(td, -7.0)
    G.knn_exact_simt(x, x, k, True, out=(ei, ed), rows=rows.to(torch.int32))
    torch.cuda.synchronize()
    r = rows.long()
    a_i, b_i = ti[r].cpu().numpy(), ei[r].cpu().numpy()
    assert np.array_equal(a_i, b_i), f"{(a_i != b_i).any(axis=1).sum()} of {len(r)} rows differ"
    assert np.array_equal(td[r].cpu().numpy().view(np.uint32), ed[r].cpu().numpy().view(np.uint32))


def _check_rows_vs_oracle(x_cpu, ti, td, k, n_rows, seed):
    rng = np.random.default_rng(seed)
    rows = np.sort(rng.choice(x_cpu.shape[0], n_rows, replace=False))
    xq = x_cpu[rows]
    # self exclusion by global index: search with the row itself admitted at k+1 and drop it
    oi, od = orc.knn_exact(xq, x_cpu, k + 1, False)
    for t, r in enumerate(rows):
        keep = oi[t] != r
        assert keep.sum() >= k
        want_i, want_d = oi[t][keep][:k], od[t][keep][:k]
        got_i, got_d = ti[r].cpu().numpy(), td[r].cpu().numpy()
        assert np.array_equal(got_i, want_i), (r, got_i, want_i)
        assert np.array_equal(got_d.view(np.uint32), want_d.view(np.uint32)), r


@pytest.mark.parametrize("modality", ["texts", "images"])
def test_c2_full_size_knn_bit_exact(modality):
    """BASELINE.json configs[1] at full size: texts 158,915 x 768 and images 31,783 x 4,096, k=15."""
    b = _bench()
    data = b.make_data(b.WORKLOADS["c2"])
    x_cpu = data[modality].numpy()
    x = data[modality].cuda()
    n = x.shape[0]
    ti, td, st = _tc(x, x, 15, True)
    assert st["rows"] == n and st["precision"] == 0, st
    assert st["fallback_rows"] == 0 and st["first_pass_uncertified"] <= 0.001 * n, st     # certified at the fp16 level
    assert (ti >= 0).all() and (ti < n).all()
    # every row against the exhaustive kernel (texts: 38.8 TFLOP of fp32 FMA work, a few seconds)
    _check_rows_vs_exact(x, ti, td, 15, torch.arange(n, device="cuda"))
    # and 256 random rows against the C oracle itself
    _check_rows_vs_oracle(x_cpu, ti, td, 15, 256, 11)


def test_c3_shaped_slice_knn_bit_exact():
    """C3 shape (BERT-like 768-D, k=15) at 200k rows: large enough for CTA pairs walking the database in
    L2 windows with the lists carried between launches."""
    b = _bench()
    wl = dict(b.WORKLOADS["c3"], mods=[("texts", 200000, 768, "bert")])
    x_cpu = b.make_data(wl)["texts"]
    x = x_cpu.cuda()
    n = x.shape[0]
    ti, td, st = _tc(x, x, 15, True)
    assert st["fallback_rows"] == 0, st
    g = torch.Generator().manual_seed(5)
    rows = torch.randperm(n, generator=g)[:40000].sort().values.cuda()
    _check_rows_vs_exact(x, ti, td, 15, rows)
    _check_rows_vs_oracle(x_cpu.numpy(), ti, td, 15, 128, 12)


def test_c4_shaped_slice_knn_bit_exact():
    """C4 shape (1,000 blobs in 128-D, k=30) at 1M rows: neighbour gaps are small against |x||y|, so the
    fp16 level cannot certify and the probe must send the data to the split-fp16 level, which must."""
    b = _bench()
    wl = dict(b.WORKLOADS["c4"], mods=[("blobs", 1000000, 128, "blobs")])
    x_cpu = b.make_data(wl)["blobs"]
    x = x_cpu.cuda()
    n = x.shape[0]
    ti, td, st = _tc(x, x, 30, True)
    assert st["precision"] == 1, st                               # level the rows were certified at
    assert st["fallback_rows"] <= 1e-4 * n, st
    g = torch.Generator().manual_seed(6)
    rows = torch.randperm(n, generator=g)[:60000].sort().values.cuda()
    _check_rows_vs_exact(x, ti, td, 30, rows)
    _check_rows_vs_oracle(x_cpu.numpy(), ti, td, 30, 128, 13)


def test_c2_spectral_init_residuals_at_full_size():
    """embed_all's contract (model.py:211-234) at BASELINE.json configs[1] size, where a dense eigensolve is out of
    reach: the 16 vectors returned for the 158,915-row text graph are unit norm, mutually orthogonal, and eigenvectors of L = I - D^-1/2 S D^-1/2 + 1e-6 I to torch.lobpcg's own tolerance
    (residual |L v - lambda v| < 3.5e-4 ... 2e-3 with fp32 operator applications), with Rayleigh quotients in [0, 1)."""
    from umap_b200 import graph as G
    from umap_b200.spectral import normalized_adjacency, spectral_init
    b = _bench()
    x = b.make_data(b.WORKLOADS["c2"])["texts"].cuda()
    idx, dist = G.knn_graph(x, x, 15, True)
    col, w, _, _ = G.smooth_knn(idx, dist, "bisect")
    g = G.fuzzy_union(col, w)
    torch.manual_seed(0)
    v = spectral_init(g, 16)
    n = g.n_rows
    assert tuple(v.shape) == (n, 16) and bool(torch.isfinite(v).all())
    v64 = v.double()
    assert torch.allclose(v64.norm(dim=0), torch.ones(16, dtype=torch.float64, device=v.device), atol=1e-4)
    gram = v64.T @ v64
    assert float((gram - torch.eye(16, dtype=torch.float64, device=v.device)).abs().max()) < 1e-3
    aval = normalized_adjacency(g)
    av = torch.zeros_like(v64)
    av.index_add_(0, g.row.long(), aval.double()[:, None] * v64[g.col.long()])
    lv = (1.0 + 1e-6) * v64 - av                                  # L v
    lam = (v64 * lv).sum(0)
    res = (lv - v64 * lam).norm(dim=0)
    assert float(res.max()) < 2e-3, res
    assert float(lam.min()) > -1e-4 and float(lam.max()) < 1.0, lam
    # (No orthogonality check against the trivial eigenvector D^1/2 1: this graph has 64 well separated clusters,
    # i.e. ~64 eigenvalues within 1e-3 of lambda_min, and embed_all drops "the first" of the computed vectors
    # (model.py:234), which in a near-degenerate cluster is an arbitrary direction of it -- for torch.lobpcg too.)
    _record("c2_texts_spectral", {"max_residual": float(res.max()), "lambda_min": float(lam.min()), "lambda_max": float(lam.max())})


# --------------------------------------------------------------------------- C1 as such
C1_REF = {"reference_200": 0.871, "newton_200": (0.834, 0.847), "bisect_200": (0.775, 0.780),
          "newton_600": (0.956, 0.957), "bisect_600": (0.960, 0.962)}


@pytest.mark.parametrize("sigma,epochs", [("newton", 200), ("bisect", 200), ("newton", 600), ("bisect", 600)])
def test_c1_fit_trustworthiness(sigma, epochs, monkeypatch):
    """BASELINE.json configs[0] exactly: 2,000 x 64 Gaussian blobs, k=15, 2-D, lr 0.01, num_rep 8, batch 256,
    through impl.util.train.  BASELINE.md: the reference's optimiser on an EXACT kNN graph reaches
    trustworthiness@15 of 0.834-0.847 (its Newton sigma) / 0.775-0.780 (bisection) at 200 epochs, where the
    run is still in its expansion transient, and 0.956-0.957 / 0.960-0.962 at the CLI default of 600 epochs
    (seed noise ~0.002); the reference as shipped (NN-descent graph) scores 0.871 at 200.  The engine must
    reach those bands from below (tolerance 0.03 at 200 epochs, where the value moves ~0.001 per epoch and depends
    on the sample stream; 0.015 at 600, measured 0.948-0.955 for the Newton sigma and 0.957-0.962 for bisection over
    three seeds); exceeding them is not a failure, the measured values are recorded (profiles/r02_quality_tests.json)."""
    from sklearn.manifold import trustworthiness
    monkeypatch.setenv("MMUMAP_SIGMA", sigma)
    b = _bench()
    util = importlib.import_module("impl.util")
    data = b.make_data(b.WORKLOADS["c1"])
    cfg = util.Config(k_neighbors=15, out_dim=2, min_dist=0.1, train_epochs=epochs, num_rep=8, lr=0.01, alpha=1.0,
                      batch_size=256, test_epochs=120)
    vals = []
    for seed in (0, 1, 2):
        torch.manual_seed(seed)
        model = util.train(data, cfg)
        assert model.encoders[0].sigma_solver == sigma
        vals.append(trustworthiness(data["blobs"].numpy(), model.embeds[0].detach().cpu().numpy(), n_neighbors=15))
    lo, hi = C1_REF[f"{sigma}_{epochs}"]
    tol = 0.03 if epochs == 200 else 0.015
    print(f"C1 {sigma} {epochs} epochs: trustworthiness@15 = {vals} (reference optimiser on the exact graph: {lo}-{hi})")
    _record(f"c1_{sigma}_{epochs}", {"trustworthiness_15": vals, "reference_band": [lo, hi]})
    # the spectral initialisation starts from a random block and C1's ten well separated blobs make the graph's lowest
    # eigenvalues a ten-fold near-degenerate cluster (any two vectors of it are a valid embed_all result, for torch.lobpcg
    # too): runs differ by a few 0.01 at 200 epochs and ~0.01 at 600.  The median of three seeds is held to the band.
    assert sorted(vals)[1] >= lo - tol, (sigma, epochs, vals, (lo, hi))
    assert min(vals) >= lo - 3 * tol, (sigma, epochs, vals, (lo, hi))


def _record(key, value):
    """measured metrics are appended to gpurun_out/quality_tests.json (copied to profiles/ by the builder)"""
    import json
    path = os.path.join(ROOT, "gpurun_out", "quality_tests.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        cur = json.load(open(path)) if os.path.exists(path) else {}
        cur[key] = value
        json.dump(cur, open(path, "w"), indent=1, sort_keys=True)
    except OSError:
        pass
