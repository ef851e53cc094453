"""ORACLE (test infrastructure, NOT product code): fp32 PyTorch-eager CPU restatement of the
reference's hot path.  **Parity unpinned by the reference**: the reference ships no tests or
golden vectors and its arithmetic lives in `diffusers` (unpinned, 0.14 <= v < 0.20,
absent from this machine), so this file restates the published diffusers-0.14 semantics of
`UNet2DConditionModel` for the SD-2.1 config, and the reference's own processor

    /root/reference/models/xti_attention_processor.py:9-57   (XTIAttenProc.__call__)

is re-executed *verbatim from the reference tree* against this file's `CrossAttention` by
tests/golden/make_golden.py to pin the cross-attention fixtures (tests/golden/xti_*.pt).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

Call sites restated:
  training/coach.py:197-198   model_pred = unet(noisy_latents, timesteps, _hs).sample
  training/coach.py:201-214   target / fp32 mse / backward
  sd_pipeline_call.py:78-101  two-pass CFG + scheduler step
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Union

import torch
import torch.nn as nn
import torch.nn.functional as F


class CrossAttention(nn.Module):
    """diffusers.models.cross_attention.CrossAttention (0.14) members used by
    xti_attention_processor.py:29-55, with SD-2.1's upcast_attention=True."""

    def __init__(self, query_dim: int, cross_attention_dim: Optional[int], heads: int, dim_head: int = 64,
                 upcast_attention: bool = True):
        super().__init__()
        inner = heads * dim_head
        kv = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads, self.scale, self.upcast_attention = heads, dim_head ** -0.5, upcast_attention
        self.cross_attention_norm = False
        self.norm_cross = None
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(kv, inner, bias=False)
        self.to_v = nn.Linear(kv, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])
        self.processor = None

    def head_to_batch_dim(self, t):
        b, n, c = t.shape
        h = self.heads
        return t.reshape(b, n, h, c // h).permute(0, 2, 1, 3).reshape(b * h, n, c // h)

    def batch_to_head_dim(self, t):
        bh, n, d = t.shape
        h = self.heads
        return t.reshape(bh // h, h, n, d).permute(0, 2, 1, 3).reshape(bh // h, n, d * h)

    def prepare_attention_mask(self, attention_mask, target_length, batch_size=None):
        return None if attention_mask is None else attention_mask

    def get_attention_scores(self, query, key, attention_mask=None):
        dtype = query.dtype
        if self.upcast_attention:
            query, key = query.float(), key.float()
        scores = torch.baddbmm(
            torch.empty(query.shape[0], query.shape[1], key.shape[1], dtype=query.dtype, device=query.device),
            query, key.transpose(-1, -2), beta=0, alpha=self.scale)
        if attention_mask is not None:
            scores = scores + attention_mask
        probs = scores.softmax(dim=-1)
        return probs.to(dtype)

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask)


class XTIAttenProcOracle:
    """Restatement of /root/reference/models/xti_attention_processor.py:9-57 (same op order)."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None):
        _ehs_bypass = None
        if encoder_hidden_states is not None:
            if isinstance(encoder_hidden_states, dict):                       # :16-22
                this_idx = encoder_hidden_states["this_idx"]
                _ehs = encoder_hidden_states[f"CONTEXT_TENSOR_{this_idx}"]
                if f"CONTEXT_TENSOR_BYPASS_{this_idx}" in encoder_hidden_states:
                    _ehs_bypass = encoder_hidden_states[f"CONTEXT_TENSOR_BYPASS_{this_idx}"]
                encoder_hidden_states["this_idx"] += 1
                encoder_hidden_states["this_idx"] %= 16
            else:
                _ehs = encoder_hidden_states                                  # :23-24
        else:
            _ehs = None                                                       # :25-26
        query = attn.to_q(hidden_states)                                      # :30
        if _ehs is None:
            _ehs = hidden_states                                              # :32-33
        key = attn.to_k(_ehs)                                                 # :38
        value = attn.to_v(_ehs_bypass if _ehs_bypass is not None else _ehs)   # :39-42
        query = attn.head_to_batch_dim(query)                                 # :44-46
        key = attn.head_to_batch_dim(key)
        value = attn.head_to_batch_dim(value)
        probs = attn.get_attention_scores(query, key, None)                   # :48
        hidden_states = torch.bmm(probs, value)                               # :49
        hidden_states = attn.batch_to_head_dim(hidden_states)                 # :50
        hidden_states = attn.to_out[0](hidden_states)                         # :53
        hidden_states = attn.to_out[1](hidden_states)                         # :55
        return hidden_states


class GEGLU(nn.Module):
    def __init__(self, c_in, c_out):
        super().__init__()
        self.proj = nn.Linear(c_in, 2 * c_out)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(c, 4 * c), nn.Dropout(0.0), nn.Linear(4 * c, c)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, c, heads, ctx_dim, ln_eps):
        super().__init__()
        self.norm1 = nn.LayerNorm(c, eps=ln_eps)
        self.attn1 = CrossAttention(c, None, heads)
        self.norm2 = nn.LayerNorm(c, eps=ln_eps)
        self.attn2 = CrossAttention(c, ctx_dim, heads)
        self.norm3 = nn.LayerNorm(c, eps=ln_eps)
        self.ff = FeedForward(c)

    def forward(self, x, ctx):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), encoder_hidden_states=ctx) + x
        x = self.ff(self.norm3(x)) + x
        return x


class Transformer2DModel(nn.Module):
    def __init__(self, c, heads, ctx_dim, groups, gn_eps, ln_eps):
        super().__init__()
        self.norm = nn.GroupNorm(groups, c, eps=gn_eps)
        self.proj_in = nn.Linear(c, c)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(c, heads, ctx_dim, ln_eps)])
        self.proj_out = nn.Linear(c, c)

    def forward(self, x, ctx):
        b, c, h, w = x.shape
        res = x
        x = self.norm(x).permute(0, 2, 3, 1).reshape(b, h * w, c)
        x = self.proj_in(x)
        for blk in self.transformer_blocks:
            x = blk(x, ctx)
        x = self.proj_out(x)
        return x.reshape(b, h, w, c).permute(0, 3, 1, 2) + res


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb, groups, eps):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Downsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class TimestepEmbedding(nn.Module):
    def __init__(self, cin, dim):
        super().__init__()
        self.linear_1 = nn.Linear(cin, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


def timestep_sinusoid(timesteps: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusers get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0)."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class _Block(nn.Module):
    pass


class _Out:
    def __init__(self, sample):
        self.sample = sample


class UNetOracle(nn.Module):
    """SD-2.1 `UNet2DConditionModel` restated; parameter names == diffusers state_dict keys."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        ch = cfg.block_out_channels
        temb = ch[0] * cfg.time_embed_dim_mult
        G, eps = cfg.norm_num_groups, cfg.norm_eps
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(ch[0], temb)

        def xf(c, heads):
            return Transformer2DModel(c, heads, cfg.cross_attention_dim, G, cfg.xf_norm_eps, cfg.ln_eps)

        self.down_blocks = nn.ModuleList()
        cin = ch[0]
        for i, cout in enumerate(ch):
            blk = _Block()
            blk.resnets = nn.ModuleList()
            blk.attentions = nn.ModuleList() if cfg.down_has_attn[i] else None
            for j in range(cfg.layers_per_block):
                blk.resnets.append(ResnetBlock2D(cin, cout, temb, G, eps))
                cin = cout
                if cfg.down_has_attn[i]:
                    blk.attentions.append(xf(cout, cfg.num_heads[i]))
            blk.downsamplers = nn.ModuleList([Downsample2D(cout)]) if i < len(ch) - 1 else None
            self.down_blocks.append(blk)
        self.mid_block = _Block()
        self.mid_block.resnets = nn.ModuleList([ResnetBlock2D(ch[-1], ch[-1], temb, G, eps) for _ in range(2)])
        self.mid_block.attentions = nn.ModuleList([xf(ch[-1], cfg.num_heads[-1])])
        self.up_blocks = nn.ModuleList()
        rev, rev_heads = list(reversed(ch)), list(reversed(cfg.num_heads))
        has_attn = list(reversed(cfg.down_has_attn))
        for i, cout in enumerate(rev):
            prev = rev[i - 1] if i > 0 else rev[0]
            inp = rev[min(i + 1, len(rev) - 1)]
            blk = _Block()
            blk.resnets = nn.ModuleList()
            blk.attentions = nn.ModuleList() if has_attn[i] else None
            n = cfg.layers_per_block + 1
            for j in range(n):
                skip = inp if j == n - 1 else cout
                hid = prev if j == 0 else cout
                blk.resnets.append(ResnetBlock2D(hid + skip, cout, temb, G, eps))
                if has_attn[i]:
                    blk.attentions.append(xf(cout, rev_heads[i]))
            blk.upsamplers = nn.ModuleList([Upsample2D(cout)]) if i < len(ch) - 1 else None
            self.up_blocks.append(blk)
        self.conv_norm_out = nn.GroupNorm(G, ch[0], eps=eps)
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, padding=1)
        self.set_attn_processor(XTIAttenProcOracle())
        self.requires_grad_(False)   # coach.py:647-648 (UNet frozen)

    def set_attn_processor(self, proc):
        for m in self.modules():
            if isinstance(m, CrossAttention):
                m.processor = proc

    def forward(self, sample, timestep, encoder_hidden_states: Union[Dict, torch.Tensor],
                cross_attention_kwargs=None):
        cfg = self.cfg
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], dtype=torch.long, device=sample.device)
        if timestep.ndim == 0:
            timestep = timestep[None]
        timestep = timestep.expand(sample.shape[0])
        temb = self.time_embedding(timestep_sinusoid(timestep, cfg.block_out_channels[0]).to(sample.dtype))
        ctx = encoder_hidden_states
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            for j, res in enumerate(blk.resnets):
                x = res(x, temb)
                if blk.attentions is not None:
                    x = blk.attentions[j](x, ctx)
                skips.append(x)
            if blk.downsamplers is not None:
                x = blk.downsamplers[0](x)
                skips.append(x)
        x = self.mid_block.resnets[0](x, temb)
        x = self.mid_block.attentions[0](x, ctx)
        x = self.mid_block.resnets[1](x, temb)
        for blk in self.up_blocks:
            for j, res in enumerate(blk.resnets):
                x = res(torch.cat([x, skips.pop()], dim=1), temb)
                if blk.attentions is not None:
                    x = blk.attentions[j](x, ctx)
            if blk.upsamplers is not None:
                x = blk.upsamplers[0](x)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        return _Out(x)


def train_step_oracle(unet: UNetOracle, noisy_latents, timesteps, target, ctx: Dict):
    """coach.py:197-214 restated: eps = unet(...).sample; loss = mse(eps.float(), target.float());
    loss.backward() -> grads on the context tensors (which must have requires_grad=True)."""
    eps = unet(noisy_latents, timesteps, ctx).sample
    loss = F.mse_loss(eps.float(), target.float(), reduction="mean")
    leaves = [v for k, v in ctx.items() if torch.is_tensor(v) and v.requires_grad] if isinstance(ctx, dict) \
        else [ctx]
    grads = torch.autograd.grad(loss, leaves, allow_unused=True)
    return eps.detach(), loss.detach(), list(grads)
