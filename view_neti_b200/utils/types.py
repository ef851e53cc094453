"""Carrier records of the conditioning path (reference utils/types.py:8-31).  Scripts written against the reference
construct them by keyword or by position, so field names and order are the reference's.

One deliberate difference: the reference declares the optional `PESigmas` fields with the TYPE `float` as their default
value (`sigma_theta: Optional[float] = float`, utils/types.py:20-23), which no consumer can use (the Fourier matrix is
scaled by these numbers, models/positional_encoding.py:166-169); here an unset bandwidth is `None`.
"""
from dataclasses import dataclass
from typing import Optional

import torch


@dataclass
class NeTIBatch:
    """One text-encoder pass of the reference: one UNet layer, B prompts (coach.py:289-296)."""
    input_ids: torch.Tensor                        # [B, 77] token ids of the prompts
    input_ids_placeholder_object: torch.Tensor     # [B] id of the object placeholder per prompt (selects the object mapper)
    input_ids_placeholder_view: torch.Tensor       # [B] id of the view placeholder per prompt, -1 when there is none
    timesteps: torch.Tensor                        # [B] diffusion timestep, first input of the mappers
    unet_layers: torch.Tensor                      # [B] index 0..15 of the cross-attention layer the pass is for
    truncation_idx: Optional[int] = None           # nested-dropout truncation (off in every shipped config)


@dataclass
class PESigmas:
    """Bandwidths of the random Fourier features in front of a NeTI mapper (config.py:142-178)."""
    sigma_t: float
    sigma_l: float
    sigma_theta: Optional[float] = None
    sigma_phi: Optional[float] = None
    sigma_r: Optional[float] = None
    sigma_dtu12: Optional[float] = None


@dataclass
class MapperOutput:
    """What a NeTI mapper returns for a batch of (timestep, layer[, view]) inputs (neti_mapper.py:416-438)."""
    word_embedding: torch.Tensor                   # [B, dim] written into the placeholder row of the token embeddings
    bypass_output: Optional[torch.Tensor]          # [B, dim] injected after the text encoder (None without output_bypass)
    bypass_unconstrained: bool                     # True: replaces the state at mean norm; False: norm-matched residual
    output_bypass_alpha: float                     # strength of the norm-matched residual bypass
