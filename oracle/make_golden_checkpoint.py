"""oracle/make_golden_checkpoint.py -- a checkpoint WRITTEN BY THE REFERENCE.

Fits the unmodified reference on a small seeded two-modality problem, saves it with its own
UMAPMixture.save_state_dict (/root/reference/impl/model.py:653-683: int64 sparse COO graphs, leaf
embeddings, CPU tensors) and records its transform of held-out rows.  tests/test_gpu_e2e.py loads
the file with the ENGINE's load_state_dict and transforms the same rows.
Build container only:   python oracle/make_golden_checkpoint.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("MMUMAP_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
from oracle.e2e_data import make_problem  # noqa: E402


def main():
    sys.path.insert(0, REF)
    from impl.util import Config, embed, train
    cfg = Config(k_neighbors=10, out_dim=4, min_dist=0.1, train_epochs=150, num_rep=8, lr=0.01, alpha=1.0,
                 batch_size=128, test_epochs=40)
    train_d, test_d = make_problem(n_train=400, n_test=60, clusters=5, seed=77)
    torch.manual_seed(5)
    model = train({k: torch.from_numpy(v) for k, v in train_d.items()}, cfg)
    out = os.path.join(ROOT, "tests", "golden", "ref_checkpoint.pt")
    model.save_state_dict(out)
    torch.manual_seed(6)
    emb = embed(model, [torch.from_numpy(test_d["texts"]), torch.from_numpy(test_d["images"])], [0, 1], cfg)
    np.savez(os.path.join(ROOT, "tests", "golden", "ref_checkpoint_transform.npz"),
             texts=emb[0].detach().numpy(), images=emb[1].detach().numpy())
    print("wrote", out, os.path.getsize(out))


if __name__ == "__main__":
    main()
