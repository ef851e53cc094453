"""Carrier dataclasses of the reference's conditioning path (reference utils/types.py:8-31), field-compatible."""
from dataclasses import dataclass
from typing import Optional

import torch


@dataclass
class NeTIBatch:
    input_ids: torch.Tensor
    input_ids_placeholder_object: torch.Tensor
    input_ids_placeholder_view: torch.Tensor
    timesteps: torch.Tensor
    unet_layers: torch.Tensor
    truncation_idx: Optional[int] = None


@dataclass
class PESigmas:
    sigma_t: float
    sigma_l: float
    sigma_theta: Optional[float] = None
    sigma_phi: Optional[float] = None
    sigma_r: Optional[float] = None
    sigma_dtu12: Optional[float] = None


@dataclass
class MapperOutput:
    word_embedding: torch.Tensor
    bypass_output: Optional[torch.Tensor]
    bypass_unconstrained: bool
    output_bypass_alpha: float
