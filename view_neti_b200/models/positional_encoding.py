"""Fourier-feature positional encoding used by the NeTI mappers (reference models/positional_encoding.py:146-195).

Only the random projection matrix lives here (the sin/cos evaluation is part of the fused mapper kernel):
`w = randn(dim // 2, nfeats)` drawn after `torch.manual_seed(seed)` on the CPU generator, column i scaled by sigmas[i]
(reference :164-169) - bit-identical to the reference for the same seed.  Unlike the reference this does NOT reseed the
global RNG as a side effect: a private generator is used.
"""
from __future__ import annotations

from typing import List

import torch


def fourier_matrix(sigmas: List[float], dim: int = 64, seed: int = 0) -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(seed)
    w = torch.randn((dim // 2, len(sigmas)), generator=g)
    for i, s in enumerate(sigmas):
        w[:, i] *= s
    return w
