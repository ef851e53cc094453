"""SD-2.1 UNet topology (config + parameter table) and seeded synthetic weights.

The reference loads `UNet2DConditionModel.from_pretrained("stabilityai/stable-diffusion-2-1",
subfolder="unet")` (reference training/coach.py:635-640).  No checkpoint exists on this
machine and there is no network, so the path runs on *seeded random weights of that exact
architecture*: same parameter names (diffusers state_dict keys), same shapes, 865.9 M params.
A real checkpoint loads through the same `state_dict` interface.

Topology facts restated from the public SD-2.1 `unet/config.json`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch


@dataclass(frozen=True)
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    sample_size: int = 96
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    # diffusers calls this "attention_head_dim" but for SD-2.1 it is the NUMBER of heads
    num_heads: Tuple[int, ...] = (5, 10, 20, 20)
    down_has_attn: Tuple[bool, ...] = (True, True, True, False)
    layers_per_block: int = 2
    cross_attention_dim: int = 1024
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    xf_norm_eps: float = 1e-6          # Transformer2DModel's GroupNorm
    ln_eps: float = 1e-5
    time_embed_dim_mult: int = 4
    context_len: int = 77

    @property
    def time_embed_dim(self) -> int:
        return self.block_out_channels[0] * self.time_embed_dim_mult

    @property
    def num_cross_layers(self) -> int:
        n_down = sum(self.layers_per_block for a in self.down_has_attn if a)
        n_up = sum(self.layers_per_block + 1 for a in self.down_has_attn if a)
        return n_down + 1 + n_up


SD21 = UNetConfig()

# A narrow UNet with the same topology (4 levels, same attention placement, head_dim 64,
# 16 cross-attention layers) for CPU-sized parity tests.
TINY = UNetConfig(block_out_channels=(64, 128, 256, 256), num_heads=(1, 2, 4, 4),
                  cross_attention_dim=128, sample_size=16)


def _resnet(prefix: str, cin: int, cout: int, temb: int) -> List[Tuple[str, Tuple[int, ...], str]]:
    p = [
        (f"{prefix}.norm1.weight", (cin,), "gamma"), (f"{prefix}.norm1.bias", (cin,), "beta"),
        (f"{prefix}.conv1.weight", (cout, cin, 3, 3), "w"), (f"{prefix}.conv1.bias", (cout,), "b"),
        (f"{prefix}.time_emb_proj.weight", (cout, temb), "w"), (f"{prefix}.time_emb_proj.bias", (cout,), "b"),
        (f"{prefix}.norm2.weight", (cout,), "gamma"), (f"{prefix}.norm2.bias", (cout,), "beta"),
        (f"{prefix}.conv2.weight", (cout, cout, 3, 3), "w_res"), (f"{prefix}.conv2.bias", (cout,), "b"),
    ]
    if cin != cout:
        p += [(f"{prefix}.conv_shortcut.weight", (cout, cin, 1, 1), "w"),
              (f"{prefix}.conv_shortcut.bias", (cout,), "b")]
    return p


def _attn(prefix: str, c: int, kv_dim: int) -> List[Tuple[str, Tuple[int, ...], str]]:
    return [
        (f"{prefix}.to_q.weight", (c, c), "w"), (f"{prefix}.to_k.weight", (c, kv_dim), "w"),
        (f"{prefix}.to_v.weight", (c, kv_dim), "w"),
        (f"{prefix}.to_out.0.weight", (c, c), "w_res"), (f"{prefix}.to_out.0.bias", (c,), "b"),
    ]


def _transformer(prefix: str, c: int, ctx: int) -> List[Tuple[str, Tuple[int, ...], str]]:
    b = f"{prefix}.transformer_blocks.0"
    p = [(f"{prefix}.norm.weight", (c,), "gamma"), (f"{prefix}.norm.bias", (c,), "beta"),
         (f"{prefix}.proj_in.weight", (c, c), "w"), (f"{prefix}.proj_in.bias", (c,), "b")]
    p += [(f"{b}.norm1.weight", (c,), "gamma"), (f"{b}.norm1.bias", (c,), "beta")]
    p += _attn(f"{b}.attn1", c, c)
    p += [(f"{b}.norm2.weight", (c,), "gamma"), (f"{b}.norm2.bias", (c,), "beta")]
    p += _attn(f"{b}.attn2", c, ctx)
    p += [(f"{b}.norm3.weight", (c,), "gamma"), (f"{b}.norm3.bias", (c,), "beta"),
          (f"{b}.ff.net.0.proj.weight", (8 * c, c), "w"), (f"{b}.ff.net.0.proj.bias", (8 * c,), "b"),
          (f"{b}.ff.net.2.weight", (c, 4 * c), "w_res"), (f"{b}.ff.net.2.bias", (c,), "b")]
    p += [(f"{prefix}.proj_out.weight", (c, c), "w_res"), (f"{prefix}.proj_out.bias", (c,), "b")]
    return p


def up_block_resnet_channels(cfg: UNetConfig, i: int, j: int) -> Tuple[int, int, int]:
    """(hidden_in, skip_in, out) channels of resnet j in up block i (diffusers get_up_block wiring)."""
    rev = list(reversed(cfg.block_out_channels))
    out = rev[i]
    prev = rev[i - 1] if i > 0 else rev[0]
    inp = rev[min(i + 1, len(rev) - 1)]
    n = cfg.layers_per_block + 1
    skip = inp if j == n - 1 else out
    hid = prev if j == 0 else out
    return hid, skip, out


def param_table(cfg: UNetConfig = SD21) -> List[Tuple[str, Tuple[int, ...], str]]:
    """[(state_dict key, shape, init kind)] in diffusers naming/order."""
    ch = cfg.block_out_channels
    temb = cfg.time_embed_dim
    p: List[Tuple[str, Tuple[int, ...], str]] = [
        ("conv_in.weight", (ch[0], cfg.in_channels, 3, 3), "w"), ("conv_in.bias", (ch[0],), "b"),
        ("time_embedding.linear_1.weight", (temb, ch[0]), "w"), ("time_embedding.linear_1.bias", (temb,), "b"),
        ("time_embedding.linear_2.weight", (temb, temb), "w"), ("time_embedding.linear_2.bias", (temb,), "b"),
    ]
    cin = ch[0]
    for i, cout in enumerate(ch):
        for j in range(cfg.layers_per_block):
            p += _resnet(f"down_blocks.{i}.resnets.{j}", cin, cout, temb)
            cin = cout
            if cfg.down_has_attn[i]:
                p += _transformer(f"down_blocks.{i}.attentions.{j}", cout, cfg.cross_attention_dim)
        if i < len(ch) - 1:
            p += [(f"down_blocks.{i}.downsamplers.0.conv.weight", (cout, cout, 3, 3), "w"),
                  (f"down_blocks.{i}.downsamplers.0.conv.bias", (cout,), "b")]
    c = ch[-1]
    p += _resnet("mid_block.resnets.0", c, c, temb)
    p += _transformer("mid_block.attentions.0", c, cfg.cross_attention_dim)
    p += _resnet("mid_block.resnets.1", c, c, temb)
    has_attn_up = list(reversed(cfg.down_has_attn))
    for i in range(len(ch)):
        for j in range(cfg.layers_per_block + 1):
            hid, skip, out = up_block_resnet_channels(cfg, i, j)
            p += _resnet(f"up_blocks.{i}.resnets.{j}", hid + skip, out, temb)
            if has_attn_up[i]:
                p += _transformer(f"up_blocks.{i}.attentions.{j}", out, cfg.cross_attention_dim)
        if i < len(ch) - 1:
            out = list(reversed(ch))[i]
            p += [(f"up_blocks.{i}.upsamplers.0.conv.weight", (out, out, 3, 3), "w"),
                  (f"up_blocks.{i}.upsamplers.0.conv.bias", (out,), "b")]
    p += [("conv_norm_out.weight", (ch[0],), "gamma"), ("conv_norm_out.bias", (ch[0],), "beta"),
          ("conv_out.weight", (cfg.out_channels, ch[0], 3, 3), "w"), ("conv_out.bias", (cfg.out_channels,), "b")]
    return p


def num_params(cfg: UNetConfig = SD21) -> int:
    return sum(math.prod(s) for _, s, _ in param_table(cfg))


def init_state_dict(cfg: UNetConfig = SD21, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Seeded variance-preserving random init (CPU generator => identical on every machine).

    "w": N(0, 1/fan_in); "w_res" (last op of every residual branch): N(0, 0.25/fan_in) so the
    residual stream grows slowly over ~60 blocks and activations stay O(1); gammas 1 +- 0.1,
    betas and biases +- 0.05, so every parameter influences the output.
    """
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, shape, kind in param_table(cfg):
        if kind in ("w", "w_res"):
            fan_in = math.prod(shape[1:])
            std = (1.0 if kind == "w" else 0.5) / math.sqrt(fan_in)
            t = torch.randn(shape, generator=g, dtype=torch.float32) * std
        elif kind == "gamma":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, dtype=torch.float32)
        else:  # "b", "beta"
            t = 0.05 * torch.randn(shape, generator=g, dtype=torch.float32)
        sd[name] = t.to(dtype)
    return sd


def cross_attn_layer_names(cfg: UNetConfig = SD21) -> List[str]:
    """Module prefixes of the 16 cross-attention layers in execution order
    (== reference constants.py:1-4 UNET_LAYERS order)."""
    names = []
    for i, a in enumerate(cfg.down_has_attn):
        if a:
            names += [f"down_blocks.{i}.attentions.{j}" for j in range(cfg.layers_per_block)]
    names.append("mid_block.attentions.0")
    for i, a in enumerate(reversed(cfg.down_has_attn)):
        if a:
            names += [f"up_blocks.{i}.attentions.{j}" for j in range(cfg.layers_per_block + 1)]
    return names
