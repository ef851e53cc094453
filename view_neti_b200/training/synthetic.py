"""Seeded synthetic instance of the COMPLETE conditioning stack at the SD-2.1 shapes (no tokenizer / checkpoints exist
offline): token + position embeddings, the 23-layer CLIP encoder, one object mapper and one (theta, phi) view mapper
(141 696 parameters each), wired as `NeTIConditioning` - what `Coach(cfg, unet, conditioning=...)` needs to
run whole train steps.  Used by bench.py (`full_step`) and scripts/full_step_bench.py."""
from __future__ import annotations

from typing import Dict

import torch

from ..models.clip_encoder import SD21_TEXT, CLIPEncoder, ClipEncoderConfig, init_state_dict
from ..models.neti_conditioning import NeTIConditioning
from ..models.neti_mapper import NeTIMapper
from ..utils.types import PESigmas

OBJECT_TOKEN_ID = 49408
VIEW_TOKEN_IDS = [49409, 49410, 49411, 49412]
VIEW_TOKENS = ["<view_0_10_1p2>", "<view_10_40_1p2>", "<view_20_70_1p2>", "<view_35_100_1p2>"]


def build_conditioning(device="cuda", cfg: ClipEncoderConfig = SD21_TEXT, seed: int = 0) -> NeTIConditioning:
    g = torch.Generator().manual_seed(seed)
    C = cfg.hidden_size
    tok = torch.randn(49408 + 8, C, generator=g) * 0.02
    pos = torch.randn(77, C, generator=g) * 0.01
    enc = CLIPEncoder(init_state_dict(cfg, seed), cfg, device)
    sig = PESigmas(sigma_t=0.03, sigma_l=2.0, sigma_theta=0.5, sigma_phi=0.5, sigma_r=0.5, sigma_dtu12=0.5)
    kw = dict(output_dim=C, arch_mlp_hidden_dims=64, arch_view_net=15, arch_view_disable_tl=False, use_nested_dropout=False,
              pe_sigmas=sig, output_bypass=True, bypass_unconstrained=True, output_bypass_alpha=0.2)
    with torch.random.fork_rng(devices=[]):          # nn.Linear init draws from the process-local default generator
        torch.manual_seed(seed + 1)
        mo = NeTIMapper(embedding_type="object", norm_scale=torch.tensor(0.3714), placeholder_object_token="<statue>", **kw).to(device)
        mv = NeTIMapper(embedding_type="view", norm_scale=torch.tensor(0.4102), placeholder_view_tokens=list(VIEW_TOKENS),
                        placeholder_view_token_ids=list(VIEW_TOKEN_IDS), **kw).to(device)
    return NeTIConditioning(tok, pos, (torch.ones(C), torch.zeros(C)), enc, {OBJECT_TOKEN_ID: mo}, mv)


def synthetic_prompt(batch: int, device="cuda", seed: int = 0) -> Dict[str, torch.Tensor]:
    """`batch` prompts of 77 token ids holding the object placeholder and one view placeholder each."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1000, 40000, (batch, 77), generator=g)
    ids[:, 0], ids[:, 5] = 49406, OBJECT_TOKEN_ID
    view = torch.tensor([VIEW_TOKEN_IDS[i % len(VIEW_TOKEN_IDS)] for i in range(batch)])
    ids[:, 3] = view
    return {"input_ids": ids.to(device), "input_ids_placeholder_object": torch.full((batch,), OBJECT_TOKEN_ID, device=device),
            "input_ids_placeholder_view": view.to(device)}
