"""TEST INFRASTRUCTURE — fp32 CPU restatement of the SD-2.1 VAE (`AutoencoderKL`) as the reference drives it.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file; the product path never does.

The reference calls the VAE at two places:
  * reference training/coach.py:165-169 — `vae.encode(pixel_values).latent_dist.sample().detach() * scaling_factor`
    every train step (frozen, no grad);
  * reference sd_pipeline_call.py:115 — `pipeline.decode_latents(latents)`: `vae.decode(latents / 0.18215).sample`,
    then `(image / 2 + 0.5).clamp(0, 1)` and NHWC float32 on the host.
The arithmetic lives in `diffusers` (unpinned, 0.14 <= v < 0.20, absent here: SURVEY.md §8c), so this is a restatement
of its published algorithm at the public `stabilityai/stable-diffusion-2-1` `vae/config.json`:
block_out_channels [128, 256, 512, 512], layers_per_block 2, latent_channels 4, norm_num_groups 32, SiLU,
GroupNorm eps 1e-6 throughout, one single-head attention block (head_dim = 512) in each mid block,
encoder downsampling = zero-pad (0,1,0,1) + 3x3 stride-2 conv, decoder upsampling = nearest x2 + 3x3 conv.
Pinning: the reference holds no test vector for this path and diffusers cannot be imported, so parity against diffusers
ITSELF is unpinned.  What pins this file instead (tests/test_vae_cpu.py): SD's VAE is the LDM / taming-transformers
autoencoder and diffusers' AutoencoderKL is a re-keyed port of it; `transformers` (installed) ships verbatim ports of that
Encoder / Decoder (ChameleonVQVAEEncoder, JanusVQVAEDecoder).  With the weights mapped the way diffusers' own LDM
conversion maps them, `encode_moments` equals the Chameleon encoder + quant_conv and `decode` equals the Janus decoder
(its three extra lowest-level attention blocks neutralised by a zero output projection) to 2e-5 at the SD-2.1 widths.
Structural pins: parameter count 83 653 863 and the diffusers-0.14 state_dict key set, so real checkpoints stay loadable.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


def _gn(sd: SD, p: str, x: Tensor, groups: int, eps: float) -> Tensor:
    return F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], eps)


def _conv(sd: SD, p: str, x: Tensor, stride: int = 1, padding: int = 1) -> Tensor:
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def resnet(sd: SD, p: str, x: Tensor, groups: int, eps: float) -> Tensor:
    """ResnetBlock2D without a time embedding: GN-SiLU-conv, GN-SiLU-conv, 1x1 shortcut when channels change."""
    h = _conv(sd, p + ".conv1", F.silu(_gn(sd, p + ".norm1", x, groups, eps)))
    h = _conv(sd, p + ".conv2", F.silu(_gn(sd, p + ".norm2", h, groups, eps)))
    if p + ".conv_shortcut.weight" in sd:
        x = _conv(sd, p + ".conv_shortcut", x, padding=0)
    return x + h


def attention(sd: SD, p: str, x: Tensor, groups: int, eps: float) -> Tensor:
    """AttentionBlock (one head over all channels): GN, q/k/v Linear, softmax(q k^T / sqrt(C)) v, proj, + residual."""
    b, c, hh, ww = x.shape
    t = _gn(sd, p + ".group_norm", x, groups, eps).reshape(b, c, hh * ww).transpose(1, 2)
    q = F.linear(t, sd[p + ".query.weight"], sd[p + ".query.bias"])
    k = F.linear(t, sd[p + ".key.weight"], sd[p + ".key.bias"])
    v = F.linear(t, sd[p + ".value.weight"], sd[p + ".value.bias"])
    s = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(c), dim=-1)
    o = F.linear(s @ v, sd[p + ".proj_attn.weight"], sd[p + ".proj_attn.bias"])
    return x + o.transpose(1, 2).reshape(b, c, hh, ww)


def _mid(sd: SD, p: str, x: Tensor, groups: int, eps: float) -> Tensor:
    x = resnet(sd, p + ".resnets.0", x, groups, eps)
    x = attention(sd, p + ".attentions.0", x, groups, eps)
    return resnet(sd, p + ".resnets.1", x, groups, eps)


def encode_moments(sd: SD, cfg, x: Tensor) -> Tuple[Tensor, Tensor]:
    """pixel_values [B,3,H,W] -> (mean, logvar) of the latent posterior, each [B,4,H/8,W/8]; logvar clamped to
    [-30, 20] as DiagonalGaussianDistribution does."""
    g, eps = cfg.norm_num_groups, cfg.norm_eps
    h = _conv(sd, "encoder.conv_in", x)
    n = len(cfg.block_out_channels)
    for i in range(n):
        for j in range(cfg.layers_per_block):
            h = resnet(sd, f"encoder.down_blocks.{i}.resnets.{j}", h, g, eps)
        if i < n - 1:
            h = _conv(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", F.pad(h, (0, 1, 0, 1)), stride=2, padding=0)
    h = _mid(sd, "encoder.mid_block", h, g, eps)
    h = _conv(sd, "encoder.conv_out", F.silu(_gn(sd, "encoder.conv_norm_out", h, g, eps)))
    m = _conv(sd, "quant_conv", h, padding=0)
    mean, logvar = m.chunk(2, dim=1)
    return mean, logvar.clamp(-30.0, 20.0)


def encode_latents(sd: SD, cfg, x: Tensor, noise: Tensor) -> Tensor:
    """coach.py:165-169 with the posterior's standard-normal draw passed in: (mean + exp(logvar/2) * noise) * 0.18215."""
    mean, logvar = encode_moments(sd, cfg, x)
    return (mean + torch.exp(0.5 * logvar) * noise) * cfg.scaling_factor


def decode(sd: SD, cfg, z: Tensor) -> Tensor:
    """latents (already divided by the scaling factor) [B,4,h,w] -> image [B,3,8h,8w]."""
    g, eps = cfg.norm_num_groups, cfg.norm_eps
    h = _conv(sd, "decoder.conv_in", _conv(sd, "post_quant_conv", z, padding=0))
    h = _mid(sd, "decoder.mid_block", h, g, eps)
    n = len(cfg.block_out_channels)
    for i in range(n):
        for j in range(cfg.layers_per_block + 1):
            h = resnet(sd, f"decoder.up_blocks.{i}.resnets.{j}", h, g, eps)
        if i < n - 1:
            h = _conv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", F.interpolate(h, scale_factor=2.0, mode="nearest"))
    return _conv(sd, "decoder.conv_out", F.silu(_gn(sd, "decoder.conv_norm_out", h, g, eps)))


def decode_latents(sd: SD, cfg, latents: Tensor) -> Tensor:
    """sd_pipeline_call.py:115 -> StableDiffusionPipeline.decode_latents: NHWC float32 image in [0, 1]."""
    img = decode(sd, cfg, latents / cfg.scaling_factor)
    return (img / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1).float()
