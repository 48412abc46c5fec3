"""End-to-end quality parity on the GPU: the engine, driven through the reference-facing API
(impl.util.train / embed), against metrics the UNMODIFIED reference produced on the same seeded
two-modality problem at the same converged horizon (tests/golden/e2e_metrics.npz, written by
oracle/make_golden_e2e.py: similarity_test and knn_test of impl/validation.py, sklearn
trustworthiness@15).  The engine's graph is exact and its sigma solver converges, so its metrics
may exceed the reference's; the assertions are one-sided with the reference's seed-to-seed
spread (+ a stated slack) as tolerance."""
import importlib
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.e2e_data import CFG, make_problem

pytestmark = pytest.mark.gpu


def similarity_metric(util, model, data, cfg):
    """restates similarity_test, /root/reference/impl/validation.py:7-38"""
    mats = [data[k] for k in data]
    embeds = util.embed(model, mats, list(range(len(mats))), cfg)
    embeds = [F.normalize(e, p=2, dim=1) for e in embeds]
    sims = [(embeds[i] * embeds[j]).sum(dim=1) for i in range(len(mats)) for j in range(i + 1, len(mats))]
    return torch.stack(sims, dim=1).mean(dim=1).mean().item()


def knn_metric(util, model, data, cfg, k):
    """restates knn_test, /root/reference/impl/validation.py:40-84 (batched instead of a row loop)"""
    mats = [data[key] for key in data]
    accs = []
    for src in range(len(mats)):
        for dst in range(src + 1, len(mats)):
            e = util.embed(model, [mats[src], mats[dst]], [src, dst], cfg)
            a, b = e[0].detach(), e[1].detach()
            d = torch.cdist(a, b)
            n = a.shape[0]
            ar = torch.arange(n, device=a.device)[:, None]
            fwd = (torch.topk(d, k, dim=1, largest=False).indices == ar).any(dim=1).sum().item()
            bwd = (torch.topk(d.T, k, dim=1, largest=False).indices == ar).any(dim=1).sum().item()
            accs.append((fwd + bwd) / (2 * n))
    return float(np.mean(accs))


@pytest.mark.parametrize("stream", ["device", "host"])
def test_fit_and_transform_quality_matches_reference(golden_dir, stream, monkeypatch):
    from sklearn.manifold import trustworthiness
    monkeypatch.setenv("MMUMAP_SAMPLE_STREAM", stream)
    util = importlib.import_module("impl.util")
    ref = np.load(os.path.join(golden_dir, "e2e_metrics.npz"))
    cfg = util.Config(**CFG)
    train_d, test_d = make_problem()
    torch.manual_seed(0)
    model = util.train({k: torch.from_numpy(v) for k, v in train_d.items()}, cfg)
    assert model.sample_stream == stream
    td = {k: torch.from_numpy(v).cuda() for k, v in test_d.items()}
    got = {
        "similarity": similarity_metric(util, model, td, cfg),
        "knn1": knn_metric(util, model, td, cfg, 1),
        "knn5": knn_metric(util, model, td, cfg, 5),
    }
    for i, name in enumerate(train_d):
        got[f"trust_{name}"] = trustworthiness(train_d[name], model.embeds[i].detach().cpu().numpy(), n_neighbors=15)
    print({k: round(v, 4) for k, v in got.items()}, {k: np.round(ref[k], 4).tolist() for k in got})
    _record(f"e2e_quality_{stream}_stream", {"engine": {k: float(v) for k, v in got.items()},
                                            "reference_runs": {k: np.asarray(ref[k], dtype=float).tolist() for k in got}})
    slack = {"similarity": 0.05, "knn1": 0.05, "knn5": 0.05, "trust_texts": 0.02, "trust_images": 0.02}
    for key, val in got.items():
        lo = float(ref[key].mean()) - 3.0 * float(ref[key].std()) - slack[key]
        assert val >= lo, f"{key}: engine {val:.4f} below reference {ref[key].mean():.4f} - tolerance ({lo:.4f})"


def test_embed_and_recon_runs_and_reconstructs(monkeypatch):
    """crossmodal path (crossmodal.py:23: embed_and_recon(model, [text], [0], [1], cfg)): the reference
    raises here; the engine returns Q x D_image reconstructions that are
    far closer to the true paired image rows than a random fitted image row is."""
    monkeypatch.setenv("MMUMAP_SAMPLE_STREAM", "device")
    util = importlib.import_module("impl.util")
    cfg = util.Config(**dict(CFG, train_epochs=200, test_epochs=40))
    train_d, test_d = make_problem()
    torch.manual_seed(0)
    model = util.train({k: torch.from_numpy(v) for k, v in train_d.items()}, cfg)
    texts = torch.from_numpy(test_d["texts"]).cuda()
    recon = util.embed_and_recon(model, [texts], [0], [1], cfg)
    assert len(recon) == 1 and tuple(recon[0].shape) == (texts.shape[0], train_d["images"].shape[1])
    r = recon[0].detach().cpu().numpy()
    assert np.all(np.isfinite(r))
    true = test_d["images"]
    err = np.linalg.norm(r - true, axis=1).mean()
    rng = np.random.default_rng(0)
    base = np.linalg.norm(train_d["images"][rng.integers(0, 1500, true.shape[0])] - true, axis=1).mean()
    assert err < 0.6 * base, (err, base)


def test_batched_metrics_equal_the_row_loop():
    """umap_b200.metrics.retrieval_accuracy (engine kNN, query mode) vs the reference's per-row loop
    (validation.py:66-78: torch.norm + topk + `idx in knns`) on random embeddings."""
    from umap_b200 import metrics
    g = torch.Generator().manual_seed(3)
    a = torch.randn(700, 8, generator=g)
    b = a + 0.8 * torch.randn(700, 8, generator=g)
    for k in (1, 5):
        correct = 0
        for i in range(a.shape[0]):
            if i in torch.topk(torch.norm(b - a[i], dim=1), k, largest=False).indices:
                correct += 1
            if i in torch.topk(torch.norm(a - b[i], dim=1), k, largest=False).indices:
                correct += 1
        want = correct / (2 * a.shape[0])
        got = metrics.retrieval_accuracy(a, b, k)
        assert abs(got - want) < 1.5 / a.shape[0], (k, got, want)       # fp32 near-ties may flip a row or two


def test_reference_written_checkpoint_loads_and_transforms(golden_dir):
    """A checkpoint written by the reference's own save_state_dict (int64 sparse COO graphs, CPU leaf
    embeddings; oracle/make_golden_checkpoint.py) loads through the engine's load_state_dict and
    serves transform(): the held-out rows land where the reference put them, up to the spread of its
    own stochastic transform (compared through nearest fitted neighbours, not coordinates)."""
    model_mod = importlib.import_module("impl.model")
    util = importlib.import_module("impl.util")
    path = os.path.join(golden_dir, "ref_checkpoint.pt")
    model = model_mod.UMAPMixture.load_state_dict(path)
    assert model.k_neighbors == 10 and model.out_dim == 4 and model.num_encoders == 2
    assert model.graphs[0].is_sparse and model.graphs[0].indices().dtype == torch.int64
    ref = np.load(os.path.join(golden_dir, "ref_checkpoint_transform.npz"))
    from oracle.e2e_data import make_problem as mp
    _, test_d = mp(n_train=400, n_test=60, clusters=5, seed=77)
    cfg = util.Config(k_neighbors=10, out_dim=4, min_dist=0.1, train_epochs=150, num_rep=8, lr=0.01, alpha=1.0,
                      batch_size=128, test_epochs=40)
    torch.manual_seed(6)
    out = util.embed(model, [torch.from_numpy(test_d["texts"]), torch.from_numpy(test_d["images"])], [0, 1], cfg)
    for m, name in enumerate(("texts", "images")):
        got = out[m].detach().cpu().numpy()
        assert got.shape == ref[name].shape and np.all(np.isfinite(got))
        fitted = model.embeds[m].detach().cpu().numpy()
        # same neighbourhood of the fitted embedding: the nearest fitted row of the engine's placement is
        # among the 15 nearest fitted rows of the reference's placement for most queries
        d_ref = np.linalg.norm(ref[name][:, None, :] - fitted[None], axis=2)
        d_got = np.linalg.norm(got[:, None, :] - fitted[None], axis=2)
        near_ref = np.argsort(d_ref, axis=1)[:, :15]
        hit = np.mean([d_got[i].argmin() in near_ref[i] for i in range(got.shape[0])])
        assert hit > 0.7, (name, hit)
    # and the engine writes a checkpoint the same loader reads back
    tmp = os.path.join(os.environ.get("TMPDIR", "/tmp"), "mmu_ckpt_roundtrip.pt")
    model.save_state_dict(tmp)
    again = model_mod.UMAPMixture.load_state_dict(tmp)
    assert torch.equal(again.embeds[0].detach().cpu(), model.embeds[0].detach().cpu())


def _record(key, value):
    """measured metrics go to gpurun_out/quality_tests.json (the builder copies it to profiles/)"""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "gpurun_out", "quality_tests.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        cur = json.load(open(path)) if os.path.exists(path) else {}
        cur[key] = value
        json.dump(cur, open(path, "w"), indent=1, sort_keys=True)
    except OSError:
        pass
