"""Test stand-in for `diffusers` (not installed, and its weights need the network): just enough of AutoencoderKL
for the reference's impl/crossmodal.py:28-36 and impl/dataset.py:6 to import and run.  decode() maps a
(N, 4, 32, 32) latent to a deterministic (N, 3, 64, 64) "image".  TEST SCAFFOLDING, not product code."""
import torch


class _Decoded:
    def __init__(self, sample):
        self.sample = sample


class AutoencoderKL:
    @classmethod
    def from_pretrained(cls, name, *args, **kwargs):
        return cls()

    def to(self, device):
        return self

    def decode(self, latent):
        up = torch.nn.functional.interpolate(latent[:, :3].float(), scale_factor=2, mode="nearest")
        return _Decoded(torch.tanh(up))
