"""ORACLE (test infrastructure, NOT product code): CPU fp32 restatement of the CLIP text-transformer encoder that the
reference's conditioning path runs 16x per step (reference models/neti_clip_text_encoder.py:53,101-108 ->
`transformers.models.clip.modeling_clip.CLIPEncoder`; the reference pins transformers==4.27.4, environment.yml:335).

The algorithm lives in that third-party dependency.  transformers IS installed here (5.5.0: same CLIPEncoderLayer
arithmetic - pre-LN, q scaled by head_dim**-0.5, additive causal mask, erf GELU for hidden_act="gelu"), so the
restatement below is PINNED against the dependency's own class: tests/test_oracle_cpu.py checks `encoder_forward` against
`hf_encoder(...)` outputs and input gradients to fp32 round-off.  Only tests/, smoke() and bench.py's CPU legs may
import this module.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F


def init_state_dict(hidden: int, heads: int, layers: int, intermediate: int, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded weights under transformers' CLIPEncoder key names.  Projections ~ N(0, 1/fan_in) with the residual
    branches (out_proj, fc2) damped by (2*layers)**-0.5 so the stream stays O(1); non-trivial biases and LayerNorm
    affine parameters so every epilogue path is exercised."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    damp = (2 * layers) ** -0.5
    for i in range(layers):
        p = f"layers.{i}."
        for n in ("q", "k", "v"):
            sd[p + f"self_attn.{n}_proj.weight"] = torch.randn(hidden, hidden, generator=g) / math.sqrt(hidden)
            sd[p + f"self_attn.{n}_proj.bias"] = 0.1 * torch.randn(hidden, generator=g)
        sd[p + "self_attn.out_proj.weight"] = torch.randn(hidden, hidden, generator=g) * damp / math.sqrt(hidden)
        sd[p + "self_attn.out_proj.bias"] = 0.05 * torch.randn(hidden, generator=g)
        sd[p + "mlp.fc1.weight"] = torch.randn(intermediate, hidden, generator=g) / math.sqrt(hidden)
        sd[p + "mlp.fc1.bias"] = 0.1 * torch.randn(intermediate, generator=g)
        sd[p + "mlp.fc2.weight"] = torch.randn(hidden, intermediate, generator=g) * damp / math.sqrt(intermediate)
        sd[p + "mlp.fc2.bias"] = 0.05 * torch.randn(hidden, generator=g)
        for n in ("layer_norm1", "layer_norm2"):
            sd[p + n + ".weight"] = 1 + 0.1 * torch.randn(hidden, generator=g)
            sd[p + n + ".bias"] = 0.05 * torch.randn(hidden, generator=g)
    return sd


def encoder_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, heads: int, layers: int, eps: float = 1e-5,
                    causal: bool = True) -> torch.Tensor:
    """transformers CLIPEncoder.forward(inputs_embeds=x, causal mask) -> last_hidden_state (before final_layer_norm),
    layer by layer as CLIPEncoderLayer.forward / CLIPAttention.forward / CLIPMLP.forward compute it."""
    nseq, L, C = x.shape
    hd = C // heads
    mask = torch.full((L, L), float("-inf"), dtype=x.dtype).triu(1) if causal else None
    for i in range(layers):
        p = f"layers.{i}."
        r = x
        h = F.layer_norm(x, (C,), sd[p + "layer_norm1.weight"], sd[p + "layer_norm1.bias"], eps)
        q = F.linear(h, sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.q_proj.bias"]) * hd ** -0.5
        k = F.linear(h, sd[p + "self_attn.k_proj.weight"], sd[p + "self_attn.k_proj.bias"])
        v = F.linear(h, sd[p + "self_attn.v_proj.weight"], sd[p + "self_attn.v_proj.bias"])
        q, k, v = (t.view(nseq, L, heads, hd).transpose(1, 2) for t in (q, k, v))
        s = q @ k.transpose(-1, -2)
        if mask is not None:
            s = s + mask
        a = torch.softmax(s, dim=-1) @ v
        a = a.transpose(1, 2).reshape(nseq, L, C)
        x = r + F.linear(a, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
        r = x
        h = F.layer_norm(x, (C,), sd[p + "layer_norm2.weight"], sd[p + "layer_norm2.bias"], eps)
        h = F.gelu(F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
        x = r + F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return x


def hf_encoder(sd: Dict[str, torch.Tensor], hidden: int, heads: int, layers: int, intermediate: int, eps: float = 1e-5):
    """The dependency's own module with these weights (eager attention), and the additive causal mask the reference's
    CLIPTextTransformer builds (neti_clip_text_encoder.py:95-98)."""
    from transformers.models.clip.modeling_clip import CLIPEncoder, CLIPTextConfig
    cfg = CLIPTextConfig(hidden_size=hidden, intermediate_size=intermediate, num_hidden_layers=layers,
                         num_attention_heads=heads, hidden_act="gelu", layer_norm_eps=eps, max_position_embeddings=77,
                         vocab_size=1000)
    cfg._attn_implementation = "eager"
    enc = CLIPEncoder(cfg).eval()
    missing = enc.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys

    def run(x: torch.Tensor) -> torch.Tensor:
        L = x.shape[1]
        mask = torch.full((L, L), float("-inf"), dtype=x.dtype).triu(1)[None, None].expand(x.shape[0], 1, L, L)
        return enc(inputs_embeds=x, attention_mask=mask)[0]

    return run
