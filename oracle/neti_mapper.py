"""ORACLE (test infrastructure, NOT product code): CPU fp32 restatement of the NeTI mapper in the paper's configuration
(arch_view_net 15) - reference models/neti_mapper.py:165-197 (forward), :542-562 (inputs scaled to [-1, 1]),
:601-608 (net), :416-438 (split word / bypass, normalise * norm_scale) and models/positional_encoding.py:174-195
(Fourier features cat[sin(W x), cos(W x)]).

PINNED: tests/golden/neti_mapper.pt holds outputs and parameter gradients of the reference's own NeTIMapper
(tests/golden/make_golden_mapper.py imports it unmodified); tests/test_oracle_cpu.py::test_mapper_oracle_matches_reference_golden
holds this restatement to them.  Only tests/, smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

NUM_UNET_LAYERS = 16


def encode_inputs(timestep: torch.Tensor, unet_layer: torch.Tensor, view_params: Optional[torch.Tensor] = None,
                  view_min: Optional[Sequence[float]] = None, view_max: Optional[Sequence[float]] = None) -> torch.Tensor:
    """neti_mapper.py:545-562: (t, l[, view parameters]) -> one row per sample in [-1, 1].  `view_params` [B, n] are the raw
    numbers parsed from the view tokens (theta, phi / phi / 12 camera entries), min-max scaled column by column
    (:294-337; a column whose min equals its max is passed through unscaled)."""
    cols = [timestep.float() / 1000 * 2 - 1, unet_layer.float() / NUM_UNET_LAYERS * 2 - 1]
    if view_params is not None:
        for j in range(view_params.shape[1]):
            v = view_params[:, j].float()
            lo, hi = float(view_min[j]), float(view_max[j])
            cols.append(v if lo == hi else (v - lo) / (hi - lo) * 2 - 1)
    return torch.stack(cols, dim=1)


def mapper_forward(state: Dict[str, torch.Tensor], w: torch.Tensor, x: torch.Tensor, norm_scale: Optional[float],
                   eps: float = 1e-5) -> Tuple[torch.Tensor, torch.Tensor]:
    """`state`: reference state_dict keys net.{0,1,3,4}.{weight,bias}, output_layer.0.{weight,bias}; `w` [32, nfeat] the
    Fourier matrix; `x` [B, nfeat] from encode_inputs.  Returns (word_embedding, bypass_output), each [B, dim]."""
    proj = x @ w.t()                                                                        # positional_encoding.py:186-189
    h = torch.cat([torch.sin(proj), torch.cos(proj)], dim=1)
    for lin, ln in (("net.0", "net.1"), ("net.3", "net.4")):                                # neti_mapper.py:603-607
        h = F.linear(h, state[lin + ".weight"], state[lin + ".bias"])
        h = F.layer_norm(h, (h.shape[1],), state[ln + ".weight"], state[ln + ".bias"], eps)
        h = F.leaky_relu(h)
    out = F.linear(h, state["output_layer.0.weight"], state["output_layer.0.bias"])         # :417
    dim = out.shape[1] // 2
    word, bypass = out[:, :dim], out[:, dim:]                                               # :426-431
    if norm_scale is not None:
        word = F.normalize(word, dim=-1) * norm_scale                                       # :434-436
    return word, bypass
