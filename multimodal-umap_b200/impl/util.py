"""Drop-in replacement for the reference's impl/util.py (ref: /root/reference/impl/util.py:6-129).

Same dataclass (nine required fields, same order) and the same four thin wrappers over
UMAPMixture.fit / transform / inverse_transform.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from .model import UMAPMixture


@dataclass
class Config:
    """ref: util.py:6-31."""
    k_neighbors: int
    out_dim: int
    min_dist: float

    train_epochs: int
    num_rep: int
    lr: float
    alpha: float
    batch_size: int

    test_epochs: int


def train(data: dict, cfg: Config) -> UMAPMixture:
    """ref: util.py:33-61.  Modalities are taken in dict order."""
    data = [data[key] for key in data]
    model = UMAPMixture(k_neighbors=cfg.k_neighbors, out_dim=cfg.out_dim, min_dist=cfg.min_dist,
                        num_encoders=len(data))
    model.fit(data, epochs=cfg.train_epochs, num_rep=cfg.num_rep, lr=cfg.lr, alpha=cfg.alpha,
              batch_size=cfg.batch_size)
    return model


def embed(model: UMAPMixture, data: list, src: list, cfg: Config) -> list:
    """ref: util.py:63-87."""
    data = [d.unsqueeze(0) if d.dim() == 1 else d for d in data]
    return model.transform(data, epochs=cfg.test_epochs, data_indices=src, num_rep=cfg.num_rep, lr=cfg.lr,
                           alpha=cfg.alpha, batch_size=cfg.batch_size)


def recon(model: UMAPMixture, embeds: list, dst: list, cfg: Config) -> list:
    """ref: util.py:89-113."""
    embeds = [e.unsqueeze(0) if e.dim() == 1 else e for e in embeds]
    return model.inverse_transform(embeds, epochs=cfg.test_epochs, data_indices=dst, num_rep=cfg.num_rep,
                                   lr=cfg.lr, alpha=cfg.alpha, batch_size=cfg.batch_size)


def embed_and_recon(model: UMAPMixture, data: list, src: list, dst: list, cfg: Config) -> list:
    """ref: util.py:115-129."""
    embeds = embed(model, data, src, cfg)
    return recon(model, embeds, dst, cfg)
