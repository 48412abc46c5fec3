"""oracle/make_golden_invert.py -- golden vectors of the reference's invert-mode loss terms.

UMAPMixture.inverse_transform raises in the reference as shipped, but
its two loss terms _inv_attr_loss / _inv_rep_loss (/root/reference/impl/model.py:336-362) are
callable on their own; this script evaluates them and their autograd gradients on seeded tensors
(including the clamp branches) and stores the results in tests/golden/invert_losses.npz.
Build container only:   python oracle/make_golden_invert.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle.make_golden import load_reference  # noqa: E402


def main():
    ref = load_reference()
    mix = ref.UMAPMixture.__new__(ref.UMAPMixture)
    g = torch.Generator().manual_seed(99)
    n_ref, q, d = 90, 40, 24
    data = torch.randn(n_ref, d, generator=g) * 1.5
    x = (data[torch.randint(0, n_ref, (q,), generator=g)] + 0.4 * torch.randn(q, d, generator=g)).requires_grad_(True)
    ii = torch.randint(0, q, (260,), generator=g)
    jj = torch.randint(0, n_ref, (260,), generator=g)
    with torch.no_grad():
        x[ii[0]] = data[jj[0]]                       # zero distance -> clamp(min=1e-6) branch
    sigma = torch.rand(n_ref, generator=g) * 2.0 + 0.1
    rho = torch.rand(n_ref, generator=g) * 3.0
    rho[jj[1]] = 100.0                               # dist - rho < 0 -> clamp branch of the repulsive term
    a, b = 1.577, 0.8951
    la = mix._inv_attr_loss(x, ii, jj, a, b, data, sigma)
    ga = torch.autograd.grad(la, x)[0]
    lr_ = mix._inv_rep_loss(x, ii, jj, data, sigma, rho)
    gr = torch.autograd.grad(lr_, x)[0]
    out = os.path.join(ROOT, "tests", "golden", "invert_losses.npz")
    np.savez(out, a=a, b=b, x=x.detach().numpy(), data=data.numpy(), ii=ii.numpy(), jj=jj.numpy(), sigma=sigma.numpy(),
             rho=rho.numpy(), attr_loss=la.item(), attr_grad=ga.numpy(), rep_loss=lr_.item(), rep_grad=gr.numpy())
    print("wrote", out, la.item(), lr_.item())


if __name__ == "__main__":
    main()
