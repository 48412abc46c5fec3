"""Seeded two-modality problem shared by oracle/make_golden_e2e.py (reference run) and
tests/test_gpu_e2e.py (engine run).  TEST INFRASTRUCTURE, NOT PRODUCT CODE."""
import numpy as np

# reference CLI defaults (main.py:13-24) except the small latent size
CFG = dict(k_neighbors=15, out_dim=8, min_dist=0.1, train_epochs=600, num_rep=8, lr=0.01, alpha=1.0,
           batch_size=256, test_epochs=120)


def make_problem(n_train: int = 1500, n_test: int = 300, clusters: int = 12, seed: int = 2024):
    """Paired rows: row i of "texts" and row i of "images" share a cluster and a latent position.
    Dict order is texts, images (dataset.py:60-63)."""
    rng = np.random.default_rng(seed)
    n = n_train + n_test
    lab = rng.integers(0, clusters, n)
    z = rng.standard_normal((clusters, 6))[lab] * 3.0 + rng.standard_normal((n, 6))      # shared latent
    wt = rng.standard_normal((6, 48)) / np.sqrt(6)
    wi = rng.standard_normal((6, 64)) / np.sqrt(6)
    texts = np.tanh(0.5 * (z @ wt) + 0.1 * rng.standard_normal((n, 48))).astype(np.float32)
    images = (2.0 * (z @ wi) + 0.5 * rng.standard_normal((n, 64))).astype(np.float32)
    train = {"texts": texts[:n_train].copy(), "images": images[:n_train].copy()}
    test = {"texts": texts[n_train:].copy(), "images": images[n_train:].copy()}
    return train, test


def paired_features(n: int, seed: int):
    """Flickr30k-SHAPED paired rows for the harness tests: texts (n x 768, tanh-bounded like BERT's pooler output,
    dataset.py:52) and images (n x 4096 = 4 x 32 x 32 SD-VAE latents, dataset.py:57-58) generated from a shared 6-D
    latent so that row i of one modality is retrievable from row i of the other.  Returns numpy arrays."""
    rng = np.random.default_rng(seed)
    lab = rng.integers(0, 10, n)
    z = rng.standard_normal((10, 6))[lab] * 3.0 + rng.standard_normal((n, 6))
    wt = np.random.default_rng(100).standard_normal((6, 768)) / np.sqrt(6)
    wi = np.random.default_rng(101).standard_normal((6, 4096)) / np.sqrt(6)
    texts = np.tanh(0.5 * (z @ wt) + 0.1 * rng.standard_normal((n, 768))).astype(np.float32)
    images = (2.0 * (z @ wi) + 0.5 * rng.standard_normal((n, 4096))).astype(np.float32)
    return {"texts": texts, "images": images}
