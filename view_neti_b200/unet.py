"""Drop-in for the object the reference calls `self.unet` / `pipeline.unet`
(diffusers `UNet2DConditionModel`, loaded at reference training/coach.py:635-640, scripts/inference.py:155-158).

Call surface kept:
    unet(sample, timestep, encoder_hidden_states, cross_attention_kwargs=None).sample     coach.py:197-198,
                                                                                          sd_pipeline_call.py:78-94
    unet.config.sample_size / unet.in_channels                                            sd_pipeline_call.py:29-30,53
    unet.set_attn_processor(XTIAttenProc())                                               coach.py:679-680
    unet.requires_grad_(False) / .train() / .eval() / .enable_gradient_checkpointing() / .to(device, dtype)
    UNet2DConditionModel.from_pretrained(path, subfolder="unet", revision=...)            coach.py:636-639

`encoder_hidden_states` follows the reference protocol (models/xti_attention_processor.py:14-26):
  dict   {"this_idx": i0, "CONTEXT_TENSOR_i": [B,77,1024], "CONTEXT_TENSOR_BYPASS_i": [B,77,1024] (optional)} —
         cross-attention layer l reads entry (i0 + l) % 16; K from CONTEXT_TENSOR, V from ..._BYPASS when present;
         the reference's counter returns to i0 after the 16 layers, so the dict is left unchanged;
  Tensor the same [B,77,1024] context for every layer, K = V source (uncond pass / original_ti).
The returned `.sample` carries an autograd node whose backward runs the dgrad-only CUDA backward and fills the
`.grad` of the context tensors exactly as `accelerator.backward(loss)` does in coach.py:214.  UNet weights are
frozen (coach.py:647-648): they never receive gradients.
"""
from __future__ import annotations

import os
from types import SimpleNamespace
from typing import Dict, List, Optional, Union

import torch

from ._abi import VNError
from .engine import UNetEngine
from .sd21 import SD21, UNetConfig, init_state_dict


class UNet2DConditionOutput:
    def __init__(self, sample: torch.Tensor):
        self.sample = sample

    def __getitem__(self, i):
        return (self.sample,)[i]


class _UNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model: "UNet2DConditionModel", sample: torch.Tensor, timestep: torch.Tensor, n_ctx: int,
                *contexts: torch.Tensor):
        nb, _, h, w = sample.shape
        plan = model.engine.plan(nb, h, w)
        plan.latents.copy_(sample)
        plan.timesteps.copy_(timestep)
        # contexts = K sources for the 16 layers followed by V sources; one fused multi-tensor copy into the plan
        dst = [plan.ctx[0, i] for i in range(n_ctx)] + [plan.ctx[1, i] for i in range(n_ctx)]
        torch._foreach_copy_(dst, [c.detach() for c in contexts])
        plan.run_forward()
        model._generation += 1
        ctx.plan, ctx.model, ctx.gen, ctx.n_ctx = plan, model, model._generation, n_ctx
        ctx.dtypes = [c.dtype for c in contexts]
        return plan.eps.to(sample.dtype, copy=True)

    @staticmethod
    def backward(ctx, d_eps: torch.Tensor):
        plan, model = ctx.plan, ctx.model
        if ctx.gen != model._generation:
            raise VNError("backward() after another forward on the same UNet: the static activation plan was "
                          "overwritten (call backward before the next forward, as coach.py:197-214 does)")
        plan.d_eps.copy_(d_eps)
        plan.run_backward()
        n = ctx.n_ctx
        needs = ctx.needs_input_grad[4:]
        g = plan.d_ctx.clone()                       # one copy out of the static plan; the grads are views of it
        grads: List[Optional[torch.Tensor]] = []
        for i in range(2 * n):
            if needs[i]:
                gi = g[i // n, i % n]
                grads.append(gi if gi.dtype == ctx.dtypes[i] else gi.to(ctx.dtypes[i]))
            else:
                grads.append(None)
        return (None, None, None, None, *grads)


class UNet2DConditionModel(torch.nn.Module):
    """Frozen SD-2.1 UNet running on the hand-written sm_100a library (no eager / CPU path)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], cfg: UNetConfig = SD21, device="cuda"):
        super().__init__()
        self.cfg = cfg
        self.config = SimpleNamespace(sample_size=cfg.sample_size, in_channels=cfg.in_channels,
                                      out_channels=cfg.out_channels, cross_attention_dim=cfg.cross_attention_dim)
        self.in_channels = cfg.in_channels
        self.engine = UNetEngine(state_dict, cfg, device)
        self.device = self.engine.dev
        self.dtype = torch.float32
        self._generation = 0
        self._attn_processor = None

    # ---- construction -------------------------------------------------------------------------------
    @classmethod
    def from_seed(cls, cfg: UNetConfig = SD21, seed: int = 0, device="cuda") -> "UNet2DConditionModel":
        """Seeded synthetic weights at the exact SD-2.1 shapes (no checkpoint can be downloaded here)."""
        return cls(init_state_dict(cfg, seed), cfg, device)

    @classmethod
    def from_pretrained(cls, path: str, subfolder: Optional[str] = None, revision: Optional[str] = None,
                        device="cuda", **_) -> "UNet2DConditionModel":
        root = os.path.join(path, subfolder) if subfolder else path
        for fn in ("diffusion_pytorch_model.bin", "diffusion_pytorch_model.pt"):
            p = os.path.join(root, fn)
            if os.path.exists(p):
                return cls(torch.load(p, map_location="cpu"), SD21, device)
        p = os.path.join(root, "diffusion_pytorch_model.safetensors")
        if os.path.exists(p):
            try:
                from safetensors.torch import load_file
            except ImportError as e:  # pragma: no cover
                raise VNError("safetensors is needed to read " + p) from e
            return cls(load_file(p), SD21, device)
        raise FileNotFoundError(f"no SD-2.1 UNet checkpoint under {root} (this machine has no network; use "
                                f"UNet2DConditionModel.from_seed for synthetic weights)")

    # ---- diffusers surface touched by the reference ----------------------------------------------------
    def set_attn_processor(self, processor) -> None:
        # XTI semantics are built into the plan; the object is kept so callers can read it back.
        self._attn_processor = processor

    def enable_gradient_checkpointing(self) -> None:
        # nothing to do: the backward already stores only what dgrad needs (coach.py:672-677 is a memory knob)
        pass

    def requires_grad_(self, requires_grad: bool = True):
        if requires_grad:
            raise VNError("the UNet is frozen on this path (coach.py:647-648); weight gradients are not computed")
        return self

    def to(self, *args, **kwargs):
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.dtype):
                self.dtype = a            # output dtype; compute is bf16 x bf16 -> fp32 regardless
        return self

    # ---- forward ---------------------------------------------------------------------------------------
    def _contexts(self, ehs: Union[Dict, torch.Tensor]) -> List[torch.Tensor]:
        n = self.cfg.num_cross_layers
        if isinstance(ehs, dict):
            i0 = int(ehs.get("this_idx", 0))
            ks, vs = [], []
            for l in range(n):
                i = (i0 + l) % n
                k = ehs[f"CONTEXT_TENSOR_{i}"]
                ks.append(k)
                vs.append(ehs.get(f"CONTEXT_TENSOR_BYPASS_{i}", k))
            return ks + vs
        if torch.is_tensor(ehs):
            return [ehs] * (2 * n)
        raise VNError("encoder_hidden_states must be a context dict or a tensor (xti_attention_processor.py:14-26)")

    def forward(self, sample: torch.Tensor, timestep, encoder_hidden_states, cross_attention_kwargs=None,
                return_dict: bool = True):
        if not sample.is_cuda:
            raise VNError("the UNet hot path has no CPU fallback: `sample` must be a CUDA tensor")
        nb = sample.shape[0]
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], dtype=torch.int64, device=sample.device)
        timestep = timestep.to(device=sample.device, dtype=torch.int64).reshape(-1).expand(nb)
        ctxs = self._contexts(encoder_hidden_states)
        for c in ctxs:
            if c.shape[0] != nb:
                raise VNError(f"context batch {c.shape[0]} != sample batch {nb}")
        out = _UNetFn.apply(self, sample, timestep, self.cfg.num_cross_layers, *ctxs)
        return UNet2DConditionOutput(out)
