"""XTIAttenProc — same name, call signature and context protocol as reference models/xti_attention_processor.py:7-57,
computed by the sm_100a library instead of torch baddbmm / softmax / bmm.

It is the per-module form of the path: usable as the attention processor of ANY module that
exposes diffusers' `CrossAttention` members (to_q, to_k, to_v, to_out[0], heads), e.g. the oracle's, with autograd
towards hidden_states and the K / V context tensors (module weights are frozen: no weight gradients).  The full UNet
engine (view_neti_b200.engine) binds the same kernels statically and does not go through this object.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .. import ops
from .._abi import VNError

BF = torch.bfloat16


def _weights(attn):
    """bf16 operand copies of the frozen projection weights, cached on the module."""
    cache = getattr(attn, "_vn_cache", None)
    key = tuple(p.data_ptr() for p in (attn.to_q.weight, attn.to_k.weight, attn.to_v.weight, attn.to_out[0].weight))
    if cache is None or cache["key"] != key:
        dev = attn.to_q.weight.device
        f = lambda w: w.detach().to(dev, BF).contiguous()            # noqa: E731
        b = lambda w: w.detach().to(dev, BF).t().contiguous()        # noqa: E731
        cache = {"key": key, "qf": f(attn.to_q.weight), "qb": b(attn.to_q.weight), "kf": f(attn.to_k.weight),
                 "kb": b(attn.to_k.weight), "vf": f(attn.to_v.weight), "vb": b(attn.to_v.weight),
                 "of": f(attn.to_out[0].weight), "ob": b(attn.to_out[0].weight),
                 "obias": attn.to_out[0].bias.detach().float().contiguous() if attn.to_out[0].bias is not None else None}
        attn._vn_cache = cache
    return cache


class _XTIAttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hidden, ctx_k, ctx_v, W, heads: int, scale: float, is_self: bool):
        B, N, C = hidden.shape
        dev = hidden.device
        inner = W["qf"].shape[0]
        hb = hidden.detach().to(BF).contiguous()
        kb_src = hb if is_self else ctx_k.detach().to(BF).contiguous()
        vb_src = hb if is_self else ctx_v.detach().to(BF).contiguous()
        L = kb_src.shape[1]
        q = torch.empty(B, N, inner, dtype=BF, device=dev)
        k = torch.empty(B, L, inner, dtype=BF, device=dev)
        v = torch.empty(B, L, inner, dtype=BF, device=dev)
        ops.gemm(hb, W["qf"], q)                                   # attn.to_q            (:30)
        ops.gemm(kb_src, W["kf"], k)                               # attn.to_k(_ehs)      (:38)
        ops.gemm(vb_src, W["vf"], v)                               # attn.to_v(bypass)    (:39-42)
        o = torch.empty(B, N, inner, dtype=BF, device=dev)
        lse = torch.empty(B, heads, N, dtype=torch.float32, device=dev)
        ops.attention_fwd(q, k, v, o, lse, heads, scale)           # :44-50, fp32 logits (upcast_attention)
        out = torch.empty(B, N, W["of"].shape[0], dtype=BF, device=dev)
        ops.gemm(o, W["of"], out, bias=W["obias"])                 # attn.to_out[0]       (:53)
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.W, ctx.heads, ctx.scale, ctx.is_self = W, heads, scale, is_self
        ctx.in_dtypes = (hidden.dtype, None if is_self else ctx_k.dtype, None if is_self else ctx_v.dtype)
        return out.to(hidden.dtype)

    @staticmethod
    def backward(ctx, d_out):
        q, k, v, o, lse = ctx.saved_tensors
        W, heads = ctx.W, ctx.heads
        B, N, inner = q.shape
        L = k.shape[1]
        dev = q.device
        dob = d_out.to(BF).contiguous()
        d_o = torch.empty_like(o)
        ops.gemm(dob, W["ob"], d_o)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        delta = torch.empty(B, heads, N, dtype=torch.float32, device=dev)
        acc = torch.zeros(2 * B * L * inner, dtype=torch.float64, device=dev) if L < 256 else None
        ops.attention_bwd(q, k, v, o, lse, d_o, delta, dq, dk, dv, heads, ctx.scale, dkv_acc=acc)
        d_hidden = torch.empty(B, N, W["qb"].shape[0], dtype=torch.float32, device=dev)
        ops.gemm(dq, W["qb"], d_hidden)
        d_k = torch.empty(B, L, W["kb"].shape[0], dtype=torch.float32, device=dev)
        d_v = torch.empty(B, L, W["vb"].shape[0], dtype=torch.float32, device=dev)
        ops.gemm(dk, W["kb"], d_k)
        ops.gemm(dv, W["vb"], d_v)
        hd, kd, vd = ctx.in_dtypes
        if ctx.is_self:
            return (d_hidden + d_k + d_v).to(hd), None, None, None, None, None, None
        return d_hidden.to(hd), d_k.to(kd), d_v.to(vd), None, None, None, None


class XTIAttenProc:

    def __call__(self, attn, hidden_states: torch.Tensor, encoder_hidden_states: Optional[Dict[str, torch.Tensor]] = None,
                 attention_mask: Optional[torch.Tensor] = None):
        _ehs_bypass = None
        if encoder_hidden_states is not None:
            if isinstance(encoder_hidden_states, dict):
                this_idx = encoder_hidden_states["this_idx"]
                _ehs = encoder_hidden_states[f"CONTEXT_TENSOR_{this_idx}"]
                if f"CONTEXT_TENSOR_BYPASS_{this_idx}" in encoder_hidden_states:
                    _ehs_bypass = encoder_hidden_states[f"CONTEXT_TENSOR_BYPASS_{this_idx}"]
                encoder_hidden_states["this_idx"] += 1
                encoder_hidden_states["this_idx"] %= 16
            else:
                _ehs = encoder_hidden_states
        else:
            _ehs = None
        if attention_mask is not None:
            raise VNError("XTIAttenProc: attention masks are not used on this path (prepare_attention_mask(None) -> None)")
        if getattr(attn, "cross_attention_norm", False):
            raise VNError("XTIAttenProc: cross_attention_norm is False for Stable Diffusion and is not implemented")
        if not hidden_states.is_cuda:
            raise VNError("XTIAttenProc runs on the CUDA library only (no CPU fallback)")
        is_self = _ehs is None
        ctx_k = hidden_states if is_self else _ehs
        ctx_v = hidden_states if is_self else (_ehs_bypass if _ehs_bypass is not None else _ehs)
        heads = attn.heads
        scale = float(getattr(attn, "scale", (attn.to_q.weight.shape[0] // heads) ** -0.5))
        if attn.to_q.weight.shape[0] // heads != 64:
            raise VNError("XTIAttenProc: the attention kernels are built for head_dim 64 (SD-2.1)")
        out = _XTIAttnFn.apply(hidden_states, ctx_k, ctx_v, _weights(attn), heads, scale, is_self)
        return attn.to_out[1](out) if len(attn.to_out) > 1 else out      # Dropout(0.0)
