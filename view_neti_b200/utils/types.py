"""Carrier records of the conditioning path.  Scripts written against the reference construct them by keyword or by
position, so names and order follow reference utils/types.py:8-31; everything else here is ours: each field is
declared once with a note on who produces and who consumes it, and the classes are generated from those tables.
"""
from dataclasses import make_dataclass
from typing import Optional

import torch

Tensor = torch.Tensor

# (name, type[, default], note)
_NETI_BATCH = [
    ("input_ids", Tensor, "token ids of the prompts, [B, 77]; read by the embedding overwrite (models/neti_conditioning.py)"),
    ("input_ids_placeholder_object", Tensor, "id of the object placeholder per prompt, [B]; selects the object mapper"),
    ("input_ids_placeholder_view", Tensor, "id of the view placeholder per prompt, [B]; -1 when the prompt has none"),
    ("timesteps", Tensor, "diffusion timestep per prompt, [B]; first input of the mappers"),
    ("unet_layers", Tensor, "index 0..15 of the cross-attention layer the pass is for, [B]; second input of the mappers"),
    ("truncation_idx", Optional[int], None, "nested-dropout truncation (off in every shipped config)"),
]
_PE_SIGMAS = [
    ("sigma_t", float, "Fourier-feature bandwidth of the timestep input"),
    ("sigma_l", float, "... of the UNet-layer input"),
    ("sigma_theta", Optional[float], None, "... of the polar camera angle (theta-phi view tokens)"),
    ("sigma_phi", Optional[float], None, "... of the azimuth"),
    ("sigma_r", Optional[float], None, "... of the camera radius (unused by arch_view_net 15)"),
    ("sigma_dtu12", Optional[float], None, "... of each of the 12 DTU camera-matrix entries"),
]
_MAPPER_OUTPUT = [
    ("word_embedding", Tensor, "[B, dim] vector written into the placeholder row of the token embeddings"),
    ("bypass_output", Optional[Tensor], "[B, dim] vector injected after the text encoder (None without output_bypass)"),
    ("bypass_unconstrained", bool, "True: the bypass replaces the state at mean norm; False: norm-matched residual"),
    ("output_bypass_alpha", float, "strength of the norm-matched residual bypass"),
]


def _record(name: str, table, doc: str):
    fields = [(f[0], f[1]) if len(f) == 3 else (f[0], f[1], f[2]) for f in table]
    cls = make_dataclass(name, fields)
    cls.__module__ = __name__
    cls.__doc__ = doc + "\n\n" + "\n".join(f"    {f[0]}: {f[-1]}" for f in table)
    return cls


NeTIBatch = _record("NeTIBatch", _NETI_BATCH, "One text-encoder pass of the reference (one UNet layer, B prompts).")
PESigmas = _record("PESigmas", _PE_SIGMAS, "Bandwidths of the random Fourier features in front of a NeTI mapper.")
MapperOutput = _record("MapperOutput", _MAPPER_OUTPUT, "What a NeTI mapper returns for a batch of (timestep, layer[, view]) inputs.")
