"""Edge cases of the reference-facing API on the GPU (the argument handling of
/root/reference/impl/util.py:33-129 and impl/model.py:483-585, 620-651)."""
import importlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _util():
    return importlib.import_module("impl.util"), importlib.import_module("impl.model")


def _data(n, d, seed, centers=4):
    rng = np.random.default_rng(seed)
    c = rng.standard_normal((centers, d)) * 4
    return torch.from_numpy((c[rng.integers(0, centers, n)] + rng.standard_normal((n, d))).astype(np.float32))


@pytest.mark.parametrize("out_dim,num_rep,stream", [(2, 8, "device"), (3, 5, "device"), (5, 8, "host"), (16, 4, "device")])
def test_single_modality_fit_transform_shapes(out_dim, num_rep, stream, monkeypatch):
    monkeypatch.setenv("MMUMAP_SAMPLE_STREAM", stream)
    util, _ = _util()
    cfg = util.Config(k_neighbors=10, out_dim=out_dim, min_dist=0.1, train_epochs=12, num_rep=num_rep, lr=0.05, alpha=0.5,
                      batch_size=4096, test_epochs=5)                      # batch_size > N: one batch
    x = _data(600, 20, 1)
    model = util.train({"only": x}, cfg)
    assert len(model.embeds) == 1 and tuple(model.embeds[0].shape) == (600, out_dim)
    assert model.embeds[0].requires_grad and model.embeds[0].is_leaf          # model.py:397,481
    assert torch.isfinite(model.embeds[0]).all()
    assert model.encoders[0].sigmas.shape == (600,) and model.encoders[0].rhos.shape == (600,)
    g = model.graphs[0]
    assert g.is_sparse and g.shape == (600, 600) and g.is_coalesced()
    out = util.embed(model, [x[7]], [0], cfg)                                # 1-D input is promoted (util.py:75)
    assert tuple(out[0].shape) == (1, out_dim) and torch.isfinite(out[0]).all()
    out = util.embed(model, [x[:33]], [0], cfg)
    assert tuple(out[0].shape) == (33, out_dim)


def test_unequal_modalities_and_subset_transform():
    util, _ = _util()
    cfg = util.Config(k_neighbors=8, out_dim=4, min_dist=0.1, train_epochs=10, num_rep=8, lr=0.05, alpha=1.0,
                      batch_size=128, test_epochs=4)
    a, b = _data(700, 12, 2), _data(450, 30, 3)                              # InfoNCE pairs min(N0, N1) rows (model.py:365)
    model = util.train({"a": a, "b": b}, cfg)
    assert [tuple(e.shape) for e in model.embeds] == [(700, 4), (450, 4)]
    out = util.embed(model, [b[:20]], [1], cfg)                              # only the second modality (data_indices)
    assert len(out) == 1 and tuple(out[0].shape) == (20, 4)
    rec = util.recon(model, [model.embeds[0].detach()[:9]], [1], cfg)        # embedding -> modality-1 data space
    assert tuple(rec[0].shape) == (9, 30) and torch.isfinite(rec[0]).all()


def test_error_behaviour():
    util, model_mod = _util()
    m = model_mod.UMAPMixture(k_neighbors=5, out_dim=2, min_dist=0.1, num_encoders=1)
    with pytest.raises(ValueError, match="Invalid mode"):                   # model.py:631-632
        m.init([_data(50, 4, 0)], mode="bogus")
    with pytest.raises(ValueError):                                          # fewer than k+1 points
        model_mod.UMAPMixture(k_neighbors=15, out_dim=2, min_dist=0.1, num_encoders=1).fit([_data(10, 4, 0)], epochs=1)
    with pytest.raises(ValueError):                                          # spectral init needs 3 (out_dim+1) points
        model_mod.UMAPMixture(k_neighbors=3, out_dim=8, min_dist=0.1, num_encoders=1).fit([_data(20, 4, 0)], epochs=1)


def test_fit_transform_returns_the_stored_embeddings():
    _, model_mod = _util()
    m = model_mod.UMAPMixture(k_neighbors=6, out_dim=2, min_dist=0.1, num_encoders=1)
    out = m.fit_transform([_data(300, 8, 5)], epochs=5, num_rep=8, lr=0.05, alpha=0.5, batch_size=64)
    assert out is m.embeds and tuple(out[0].shape) == (300, 2)               # model.py:510-525
