"""Synthetic ERA-Interim-shaped fields.

Stands in for the reference's NetCDF pipeline
(/root/reference/DoWnGAN/helpers/gen_experiment_datasets.py:195-209: every
field standardised to zero mean / unit std, except the land-sea mask which is
left as 0/1).  Host-side, torch CPU RNG only, so the same call gives the same
bytes in the build container and on the GPU box.
"""
from __future__ import annotations

import torch


def synth_batch(b: int, cin: int, hc: int, up: int = 8, npred: int = 2, seed: int = 1234, aseed: int = 4321):
    """Returns (coarse (b,cin,hc,hc), fine (b,npred,hc*up,hc*up), alpha (b,1,1,1)), all fp32 on CPU."""
    g = torch.Generator().manual_seed(seed)
    coarse = torch.randn(b, cin, hc, hc, generator=g)
    if cin >= 3:  # covariate channel 2 = land-sea mask, unstandardised 0/1
        coarse[:, 2] = (torch.rand(b, hc, hc, generator=g) > 0.5).float()
    fine = torch.randn(b, npred, hc * up, hc * up, generator=g)
    alpha = torch.rand(b, 1, 1, 1, generator=torch.Generator().manual_seed(aseed))
    return coarse, fine, alpha
